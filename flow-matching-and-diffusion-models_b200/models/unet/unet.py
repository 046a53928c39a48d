"""EfficientUNetND — the CompVis/OpenAI-style denoiser (`src/models/unet/unet.py:42-326`) on the B200 blocks.

Same constructor arguments, children (`time_embed`, `input_blocks`, `middle_block`, `output_blocks`, `out`) and
state_dict keys as the reference.  ResBlocks use scale-shift conditioning (GroupNorm * (1+scale) + shift fused in
K2); the decoder's `torch.cat([h, hs.pop()])` (`unet.py:321-322`) is passed down as a virtual concat."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import ParamCache, f32, out_of_scope
from ...nn.blocks.attention import ContextBlock, SpatialCrossAttention, SpatialSelfAttention
from ...nn.blocks.residual import ResBlockND, zero_module
from ...nn.blocks.timestep import TimestepBlock
from ...nn.ops.convolution import ConvND
from ...nn.ops.normalization import fused_group_norm, fused_group_norm_table, make_group_norm
from ...nn.ops.upsampling import DownsampleND, UpsampleND
from .base import BaseUNetND
from .utils import build_timestep_features, time_mlp


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Sequential that routes `emb` to TimestepBlocks and `context` to ContextBlocks (`unet.py:18-39`)."""

    def forward(self, x, emb: torch.Tensor, context: Optional[torch.Tensor] = None):
        layers = list(self)
        i = 0
        while i < len(layers):
            layer = layers[i]
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            first = x[0] if isinstance(x, (tuple, list)) else x
            if (isinstance(layer, ResBlockND) and isinstance(nxt, UpsampleND) and nxt.can_fuse_into_producer()
                    and first.is_cuda and layer._fast_ok()):
                # the upsampler's nearest-2x is folded into the store of the ResBlock's last conv
                x = nxt(layer(x, emb, upsample_out=True), upsampled=True)
                i += 2
                continue
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, ContextBlock):
                x = layer(x, context)
            else:
                x = layer(x)
            i += 1
        return x


class EfficientUNetND(BaseUNetND):
    def __init__(self, spatial_dims: int, in_channels: int, model_channels: int, out_channels: int,
                 num_res_blocks: int, attention_resolutions: Sequence[int], dropout: float = 0.0,
                 channel_mult: Tuple[int, ...] = (1, 2, 3, 4), conv_resample: bool = True, dim_head: int = 64,
                 num_heads: int = 4, use_linear_attn: bool = True, use_scale_shift_norm: bool = True,
                 pool_factor: int = 1, cross_attention_resolutions: Optional[Sequence[int]] = None,
                 cross_attention_dim: int = 4, cross_attention_in_middle: bool = False,
                 emb_activation_before_proj: bool = False):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        if pool_factor > 1:
            raise NotImplementedError("fmdm_b200: EfficientUNetND(pool_factor>1) (PoolND/UnPoolND patchify) is "
                                      "outside the B200 hot path (SURVEY.md §2 row 1, out of scope)")
        self.spatial_dims, self.in_channels, self.model_channels = spatial_dims, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.cross_attention_resolutions = tuple(cross_attention_resolutions or ())
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_heads, self.pool_factor = num_heads, pool_factor
        self.cross_attention_dim, self.cross_attention_in_middle = cross_attention_dim, cross_attention_in_middle
        self.emb_activation_before_proj = emb_activation_before_proj

        temb_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, temb_dim), nn.SiLU(), nn.Linear(temb_dim, temb_dim))
        self.pool = nn.Identity()

        def res(cin, cout=None):
            return ResBlockND(spatial_dims=spatial_dims, channels=cin, emb_channels=temb_dim, out_channels=cout,
                              dropout=dropout, use_scale_shift_norm=use_scale_shift_norm,
                              emb_activation_before_proj=emb_activation_before_proj)

        def attn_layers(ch, ds, linear):
            layers = []
            if ds in self.attention_resolutions:
                layers.append(SpatialSelfAttention(dim=ch, heads=num_heads, dim_head=dim_head, use_linear=linear,
                                                   use_efficient_attn=True))
            if ds in self.cross_attention_resolutions:
                layers.append(SpatialCrossAttention(dim=ch, context_dim=cross_attention_dim, heads=num_heads,
                                                    dim_head=dim_head, use_linear=linear, use_efficient_attn=True))
            return layers

        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(ConvND(spatial_dims, in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                layers += attn_layers(ch, ds, use_linear_attn)
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(DownsampleND(spatial_dims, ch, use_conv=conv_resample)))
                skip_chans.append(ch)
                ds *= 2

        mid = [res(ch), SpatialSelfAttention(ch, heads=num_heads, dim_head=dim_head, use_linear=False,
                                             use_efficient_attn=True)]
        if self.cross_attention_in_middle or ds in self.cross_attention_resolutions:
            mid.append(SpatialCrossAttention(dim=ch, context_dim=cross_attention_dim, heads=num_heads,
                                             dim_head=dim_head, use_linear=False, use_efficient_attn=True))
        mid.append(res(ch))
        self.middle_block = TimestepEmbedSequential(*mid)

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                layers += attn_layers(ch, ds, use_linear_attn)
                if level and i == num_res_blocks:
                    layers.append(UpsampleND(spatial_dims, ch, use_conv=conv_resample))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(make_group_norm(ch, groups=32), nn.SiLU(),
                                 zero_module(ConvND(spatial_dims, model_channels, out_channels, 3, padding=1)))
        self.unpool = nn.Identity()
        self._cache = ParamCache()
        self.set_weight_split(model_channels * max(channel_mult) <= self.SPLIT_WEIGHT_MAX_CHANNELS)

    # ------------------------------------------------------------------------------------------------------
    def _prepare_input(self, x, context, context_ca):
        if context_ca is not None and not (self.cross_attention_resolutions or self.cross_attention_in_middle):
            raise ValueError("context_ca provided but cross-attention is disabled.")
        return (x, context)

    def _build_time_embedding(self, t, x: torch.Tensor, *, t_table=None, step_dev=None) -> torch.Tensor:
        feats = build_timestep_features(t, self.model_channels, flip_sin_to_cos=False, freq_shift=0,
                                        batch=x.shape[0], t_table=t_table, step_dev=step_dev)
        return time_mlp(feats, self.time_embed[0], self.time_embed[2])

    def _run_network(self, x, emb: torch.Tensor, context_ca) -> torch.Tensor:
        x, context = x
        ops.require_cuda(x, "EfficientUNetND.forward")
        emb = self._pack_temb(emb)
        if self.spatial_dims != 2:
            out_of_scope("EfficientUNetND with spatial_dims != 2")
        stem = self.input_blocks[0][0].conv
        cin = x.shape[1] + (context.shape[1] if context is not None else 0)
        if cin != stem.in_channels:
            raise ValueError(f"EfficientUNetND expected {stem.in_channels} input channels, got {cin}")
        if cin <= 8:
            packed = None if self.weight_split else self._cache.get("stem_tc", [stem.weight],
                                                                    lambda: ops.stem_pack(stem.weight))
            h = ops.conv_stem(x, context, f32(stem.weight), f32(stem.bias), packed=packed)
        else:
            h = self.input_blocks[0][0](x if context is None else torch.cat([x, context.to(x.dtype)], 1))
        hs = [h]
        for block in list(self.input_blocks)[1:]:
            h = block(h, emb, context_ca)
            hs.append(h)
        h = self.middle_block(h, emb, context_ca)
        for block in self.output_blocks:
            h = block((h, hs.pop()), emb, context_ca)
        head = self.out[2].conv
        if head.out_channels <= 4:
            tab = fused_group_norm_table(self.out[0], [h], silu=True)
            if tab is not None:
                return ops.conv_head(h, f32(head.weight), f32(head.bias), norm=tab)
        h = fused_group_norm(self.out[0], [h], silu=True)
        if head.out_channels <= 4:
            return ops.conv_head(h, f32(head.weight), f32(head.bias))
        return self.out[2](h).float().contiguous()
