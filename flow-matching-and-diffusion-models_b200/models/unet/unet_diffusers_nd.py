"""UNetDiffusersND — the diffusers-UNet2DModel-compatible denoiser (`src/models/unet/unet_diffusers_nd.py:19-191`)
assembled from the B200 blocks.  Constructor arguments, attribute names and state_dict keys are the reference's.

What differs is the execution: the `torch.cat([x, context])` of the sampling loop and `_prepare_input`, the optional
2x-1 centering, the fp32->bf16 cast and the NCHW->NHWC layout change all happen inside the stem kernel; activations
stay bf16 NHWC between kernels; skip concats are virtual; the head conv emits the fp32 NCHW prediction."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import f32, out_of_scope
from ...nn.blocks import DownBlock2DCompat, UNetMidBlock2DCompat, UpBlock2DCompat
from ...nn.ops.convolution import ConvND
from ...nn.ops.normalization import fused_group_norm, fused_group_norm_table, make_group_norm
from .base import BaseUNetND
from .utils import TimestepEmbedding, build_timestep_features

_DOWN = {"DownBlock2D", "AttnDownBlock2D", "CrossAttnDownBlock2D"}
_UP = {"UpBlock2D", "AttnUpBlock2D", "CrossAttnUpBlock2D"}


class UNetDiffusersND(BaseUNetND):
    def __init__(self, spatial_dims: int = 2, sample_size: int | Sequence[int] | None = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types: Sequence[str] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D",
                                                    "AttnDownBlock2D"),
                 mid_block_type: str | None = "UNetMidBlock2D",
                 up_block_types: Sequence[str] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Sequence[int] = (224, 448, 672, 896), layers_per_block: int = 2,
                 downsample_padding: int = 1, dropout: float = 0.0, attention_head_dim: int = 8,
                 norm_num_groups: int = 32, norm_eps: float = 1e-5, resnet_time_scale_shift: str = "default",
                 add_attention: bool = True, cross_attention_dim: int | None = None, **_kwargs):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.center_input_sample = center_input_sample
        self.sample_size = sample_size
        self.time_embedding_type = time_embedding_type
        self.flip_sin_to_cos = flip_sin_to_cos
        self.freq_shift = freq_shift
        self.block_out_channels = tuple(block_out_channels)
        self.cross_attention_dim = int(cross_attention_dim) if cross_attention_dim is not None else None
        chans = self.block_out_channels
        temb_dim = chans[0] * 4

        self.conv_in = ConvND(spatial_dims, in_channels, chans[0], kernel_size=3, padding=1).conv
        self.time_proj_dim = chans[0]
        self.time_embedding = TimestepEmbedding(self.time_proj_dim, temb_dim)
        self.class_embedding = None
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        common = dict(spatial_dims=spatial_dims, temb_channels=temb_dim, eps=norm_eps, groups=norm_num_groups,
                      dropout=dropout, time_scale_shift=resnet_time_scale_shift,
                      attention_head_dim=attention_head_dim)

        prev = chans[0]
        for i, kind in enumerate(down_block_types):
            if kind not in _DOWN:
                raise ValueError(f"Unsupported down block type in compat model: {kind}")
            self.down_blocks.append(DownBlock2DCompat(
                num_layers=layers_per_block, in_channels=prev, out_channels=chans[i],
                add_downsample=(i != len(chans) - 1), with_attention=kind != "DownBlock2D",
                cross_attention_dim=self.cross_attention_dim if kind == "CrossAttnDownBlock2D" else None, **common))
            prev = chans[i]

        self.mid_block = None if mid_block_type is None else UNetMidBlock2DCompat(
            in_channels=chans[-1], add_attention=add_attention,
            cross_attention_dim=self.cross_attention_dim if mid_block_type == "UNetMidBlock2DCrossAttn" else None,
            **common)

        rev = chans[::-1]
        prev = rev[0]
        for i, kind in enumerate(up_block_types):
            if kind not in _UP:
                raise ValueError(f"Unsupported up block type in compat model: {kind}")
            self.up_blocks.append(UpBlock2DCompat(
                num_layers=layers_per_block + 1, in_channels=rev[min(i + 1, len(chans) - 1)], out_channels=rev[i],
                prev_output_channel=prev, add_upsample=(i != len(chans) - 1), with_attention=kind != "UpBlock2D",
                cross_attention_dim=self.cross_attention_dim if kind == "CrossAttnUpBlock2D" else None, **common))
            prev = rev[i]

        self.conv_norm_out = make_group_norm(chans[0], groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = ConvND(spatial_dims, chans[0], out_channels, kernel_size=3, padding=1).conv
        self.set_weight_split(max(chans) <= self.SPLIT_WEIGHT_MAX_CHANNELS)

    # ------------------------------------------------------------------------------------------------------
    def _prepare_input(self, x, context=None, context_ca=None):
        # the concat / centering are folded into the stem kernel: keep the pieces apart
        return (x, context)

    def _build_time_embedding(self, t, x: torch.Tensor, *, t_table=None, step_dev=None) -> torch.Tensor:
        if self.time_embedding_type != "positional":
            raise ValueError("UNetDiffusersND currently supports positional time embedding only for strict compat.")
        feats = build_timestep_features(t, self.time_proj_dim, max_period=10000,
                                        flip_sin_to_cos=self.flip_sin_to_cos, freq_shift=self.freq_shift,
                                        batch=x.shape[0], t_table=t_table, step_dev=step_dev)
        return self.time_embedding(feats)

    def _stem(self, x: torch.Tensor, context) -> torch.Tensor:
        ops.require_cuda(x, "UNetDiffusersND.forward")
        scale, shift = (2.0, -1.0) if self.center_input_sample else (1.0, 0.0)
        cin = x.shape[1] + (context.shape[1] if context is not None else 0)
        if cin != self.conv_in.in_channels:
            raise ValueError(f"UNetDiffusersND expected {self.conv_in.in_channels} input channels, got {cin}")
        if self.spatial_dims != 2:
            out_of_scope("UNetDiffusersND with spatial_dims != 2")
        if cin <= 8:
            w = self.conv_in.weight
            # split-weight (narrow) models keep the fp32 CUDA-core stem; the others take the tensor-core stem when large
            packed = None if self.weight_split else self._stem_cache_get("stem_tc", [w], lambda: ops.stem_pack(w))
            return ops.conv_stem(x, context, f32(w), f32(self.conv_in.bias), in_scale=scale, in_shift=shift,
                                 packed=packed)
        full = x if context is None else torch.cat([x, context.to(x.dtype)], 1)
        if self.center_input_sample:
            full = 2 * full - 1.0
        if cin % 8:
            out_of_scope(f"stem conv with {cin} input channels")
        pw = self._stem_pack()
        return ops.conv2d([ops.to_nhwc_bf16(full)], pw, bias=f32(self.conv_in.bias))

    def _stem_cache_get(self, key, params, build):
        if not hasattr(self, "_stem_cache"):
            from ..._runtime import ParamCache

            self._stem_cache = ParamCache()
        return self._stem_cache.get(key, params, build)

    def _stem_pack(self):
        w = self.conv_in.weight
        return self._stem_cache_get("stem", [w], lambda: ops.pack_conv_weight([(w, 0, w.shape[1])]))

    def _run_network(self, x, emb: torch.Tensor, context_ca) -> torch.Tensor:
        x, context = x
        emb = self._pack_temb(emb)
        sample = self._stem(x, context)
        skips = [sample]
        for block in self.down_blocks:
            sample, outs = block(sample, emb, context=context_ca)
            skips.extend(outs)
        if self.mid_block is not None:
            sample = self.mid_block(sample, emb, context=context_ca)
        for block in self.up_blocks:
            n = len(block.resnets)
            res, skips = skips[-n:], skips[:-n]
            sample = block(sample, tuple(res), emb, context=context_ca)
        if self.conv_out.out_channels <= 4:
            # conv_norm_out + SiLU ride in the head kernel's load path (no normalised tensor in HBM)
            tab = fused_group_norm_table(self.conv_norm_out, [sample], silu=True)
            if tab is not None:
                return ops.conv_head(sample, f32(self.conv_out.weight), f32(self.conv_out.bias), norm=tab)
        sample = fused_group_norm(self.conv_norm_out, [sample], silu=True)
        return self._head(sample)

    def _head(self, sample: torch.Tensor) -> torch.Tensor:
        co = self.conv_out
        if co.out_channels <= 4:
            return ops.conv_head(sample, f32(co.weight), f32(co.bias))
        if co.out_channels % 8:
            out_of_scope(f"head conv with {co.out_channels} output channels")
        if not hasattr(self, "_head_cache"):
            from ..._runtime import ParamCache

            self._head_cache = ParamCache()
        pw = self._head_cache.get("head", [co.weight], lambda: ops.pack_conv_weight([(co.weight, 0, co.in_channels)]))
        return ops.conv2d([sample], pw, bias=f32(co.bias)).float().contiguous()


UNetExactND = UNetDiffusersND
