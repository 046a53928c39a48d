"""ResBlockND — drop-in for `src/nn/blocks/residual.py:13-140` running on the fused B200 kernels.

Kernel schedule of one forward (reference lines in brackets):
  K2  GroupNorm1 + SiLU over the (virtually concatenated) input            [residual.py:95-96]
  lin temb projection SiLU -> Linear, conv1's bias folded in                [residual.py:99-108]
  K1  conv1 (3x3) with the per-sample temb vector added in the epilogue    [residual.py:97, 111-112]
  K2  GroupNorm2 (+ scale-shift) + SiLU                                    [residual.py:113-116]
  K1  conv2 (3x3) with the skip path fused: identity skip = residual read in the epilogue, 1x1/3x3 conv skip =
      extra K segments accumulated into the same TMEM tile                 [residual.py:118-120]
Children and state_dict keys are identical to the reference (norm1, conv1.conv, emb_layers, norm2, conv2.conv,
skip_connection.conv)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import ParamCache, f32, out_of_scope
from ..ops.convolution import ConvND
from ..ops.normalization import RMSNormND, fused_group_norm, fused_group_norm_table, make_group_norm
from .common import zero_module
from .timestep import TimestepBlock


class TembPack:
    """All time-embedding projections of a model in ONE launch.

    The reference projects `temb` separately inside each of its ~35 ResBlocks (`residual.py:99-108`); the input is
    the same for all of them, so the denoiser concatenates every block's `emb_layers` weight into one
    [sum(out), emb] matrix (conv1's bias folded in for the add-to-hidden variant) and runs one small fp32 kernel
    per step.  Blocks receive this object in place of `emb` and take their row slice."""

    def __init__(self, raw: torch.Tensor, proj: torch.Tensor, offsets: dict):
        self.raw, self.proj, self.offsets = raw, proj, offsets

    def has(self, block) -> bool:
        return id(block) in self.offsets

    def slice(self, block) -> torch.Tensor:
        lo, n = self.offsets[id(block)]
        return self.proj[:, lo:lo + n]

    @staticmethod
    def plan(blocks):
        """(weight [sumO][I], bias [sumO], offsets) for the blocks that can use the batched projection."""
        ws, bs, offsets, lo = [], [], {}, 0
        silu = None
        for blk in blocks:
            if not (blk.uses_embedding and (blk.use_scale_shift_norm or blk.add_embedding_to_hidden)):
                continue
            if silu is None:
                silu = blk.emb_activation_before_proj
            if blk.emb_activation_before_proj != silu:
                continue
            w = blk.emb_layers.weight.detach().float()
            b = blk.emb_layers.bias.detach().float()
            if not blk.use_scale_shift_norm and blk.conv1.conv.bias is not None:
                b = b + blk.conv1.conv.bias.detach().float()
            n = w.shape[0]
            pad = (-n) % 4  # keep every slice 16-byte aligned
            ws.append(w)
            bs.append(b)
            if pad:
                ws.append(torch.zeros((pad, w.shape[1]), dtype=w.dtype, device=w.device))
                bs.append(torch.zeros((pad,), dtype=b.dtype, device=b.device))
            offsets[id(blk)] = (lo, n)
            lo += n + pad
        if not ws:
            return None
        return torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous(), offsets, bool(silu)


class ResBlockND(TimestepBlock):
    weight_split = False  # split-bf16 weights for the fused conv2 + skip matrix (see ConvND.weight_split)

    def __init__(self, channels: int, emb_channels: Optional[int], dropout: float, out_channels: int = None,
                 use_conv: bool = False, use_scale_shift_norm: bool = False, spatial_dims: int = 2,
                 norm_type: str = "gn", act: str = "silu", norm_groups: int = 32, norm_eps: float = 1e-5,
                 zero_init_last_conv: bool = True, emb_activation_before_proj: bool = False,
                 add_embedding_to_hidden: bool = False):
        super().__init__()
        if emb_channels is None and use_scale_shift_norm:
            raise ValueError("use_scale_shift_norm requires emb_channels to be provided.")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.uses_embedding = emb_channels is not None
        self.use_scale_shift_norm = bool(use_scale_shift_norm and self.uses_embedding)
        self.emb_activation_before_proj = emb_activation_before_proj
        self.add_embedding_to_hidden = add_embedding_to_hidden
        self.spatial_dims = spatial_dims
        self.norm_type = norm_type.lower()
        self.act_name = act.lower()

        oc = self.out_channels
        self.norm1 = self._make_norm(norm_type, channels, norm_groups, norm_eps)
        self.act1 = self._make_act(act)
        self.conv1 = ConvND(spatial_dims, channels, oc, 3, padding=1)
        if self.uses_embedding:
            self.emb_act = self._make_act(act)
            self.emb_layers = nn.Linear(emb_channels, 2 * oc if self.use_scale_shift_norm else oc)
        else:
            self.emb_layers = None
        self.norm2 = self._make_norm(norm_type, oc, norm_groups, norm_eps)
        self.act2 = self._make_act(act)
        self.dropout_layer = nn.Dropout(p=dropout)
        self.conv2 = ConvND(spatial_dims, oc, oc, 3, padding=1)
        if zero_init_last_conv:
            self.conv2 = zero_module(self.conv2)
        if oc == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = ConvND(spatial_dims, channels, oc, 3, padding=1)
        else:
            self.skip_connection = ConvND(spatial_dims, channels, oc, 1)
        self._cache = ParamCache()

    # ------------------------------------------------------------------------------------------------------
    def _fast_ok(self) -> bool:
        return (self.spatial_dims == 2 and self.norm_type == "gn" and self.act_name in ("silu", "swish")
                and self.conv1.fast_path_ok() and self.conv2.fast_path_ok()
                and not (self.training and self.dropout > 0))

    def _fused_tail_weight(self, split):
        """conv2's 3x3 weights followed by the skip conv's K segments, one packed K-major matrix."""
        skip = self.skip_connection
        w2, ws = self.conv2.conv.weight, skip.conv.weight

        def build():
            parts = [(w2, 0, self.out_channels)]
            c0 = 0
            for c in split:
                parts.append((ws, c0, c))
                c0 += c
            pw = ops.pack_conv_weight(parts, split=self.weight_split)
            bias = f32(self.conv2.conv.bias)
            if skip.conv.bias is not None:
                bias = bias + f32(skip.conv.bias) if bias is not None else f32(skip.conv.bias)
            return pw, bias

        key = f"tail{int(self.weight_split)}:" + ",".join(map(str, split))
        return self._cache.get(key, [w2, ws, self.conv2.conv.bias, skip.conv.bias], build)

    def forward(self, x, emb: Optional[torch.Tensor] = None, upsample_out: bool = False) -> torch.Tensor:
        """x: (N, C, H, W) tensor, or a tuple of tensors read as their channel concat (never materialised).
        upsample_out (fast path only): return the result nearest-2x upsampled (folded into conv2's store)."""
        srcs = list(x) if isinstance(x, (tuple, list)) else [x]
        if not self._fast_ok() or not srcs[0].is_cuda:
            ops.require_cuda(srcs[0], "ResBlockND.forward")
            out_of_scope(f"ResBlockND(spatial_dims={self.spatial_dims}, norm={self.norm_type}, "
                         f"act={self.act_name}, training dropout={self.dropout})")
        srcs = [ops.to_nhwc_bf16(s) for s in srcs]
        if sum(s.shape[1] for s in srcs) != self.channels:
            raise ValueError(f"ResBlockND expected {self.channels} input channels")
        if len(srcs) > 2:
            srcs = [ops.to_nhwc_bf16(torch.cat(srcs, 1))]

        # GroupNorm apply + SiLU fold into the consumer conv's operand path wherever the conv runs in row mode
        # (image rows of >= 65 pixels): the normalised tensor is never written.  Elsewhere: the standalone K2 kernel.
        _, _, hh, ww = srcs[0].shape
        split = [s.shape[1] for s in srcs]
        tab1 = None
        if ops.conv_operand_norm_ok(hh, ww, 1, [3] * len(srcs), split):
            tab1 = fused_group_norm_table(self.norm1, srcs, silu=True)
        if tab1 is None:
            h = fused_group_norm(self.norm1, srcs, silu=True)

        addvec, scale_shift, bias1 = None, None, f32(self.conv1.conv.bias)
        if self.uses_embedding and isinstance(emb, TembPack) and emb.has(self):
            # projection already computed for every block of the model in one batched launch
            if self.use_scale_shift_norm:
                scale_shift = emb.slice(self)
            elif self.add_embedding_to_hidden:
                addvec, bias1 = emb.slice(self), None  # conv1's bias is folded into the packed bias
        elif self.uses_embedding:
            if isinstance(emb, TembPack):
                emb = emb.raw
            if emb is None:
                raise ValueError("ResBlockND expects `emb` when emb_channels is set.")
            if self.use_scale_shift_norm:
                scale_shift = ops.linear_f32(emb, f32(self.emb_layers.weight), f32(self.emb_layers.bias),
                                             silu_in=self.emb_activation_before_proj)
            elif self.add_embedding_to_hidden:
                addvec = ops.linear_f32(emb, f32(self.emb_layers.weight), f32(self.emb_layers.bias), bias1,
                                        silu_in=self.emb_activation_before_proj)
                bias1 = None
        if tab1 is not None:
            offs = [sum(split[:i]) for i in range(len(split))]
            h = ops.conv2d(srcs, self.conv1.packed(split), bias=bias1, addvec=addvec, want_stats=True,
                           norm=[(tab1, o) for o in offs])
        else:
            h = ops.conv2d([h], self.conv1.packed([self.channels]), bias=bias1, addvec=addvec, want_stats=True)
        tab2 = None
        if ops.conv_operand_norm_ok(hh, ww, 1, [3], [self.out_channels]):
            tab2 = fused_group_norm_table(self.norm2, [h], silu=True, scale_shift=scale_shift)
        if tab2 is None:
            h = fused_group_norm(self.norm2, [h], silu=True, scale_shift=scale_shift)
        n2 = None if tab2 is None else [(tab2, 0)]

        if isinstance(self.skip_connection, nn.Identity):
            if len(srcs) != 1:
                srcs = [ops.to_nhwc_bf16(torch.cat(srcs, 1))]
            return ops.conv2d([h], self.conv2.packed([self.out_channels]), bias=f32(self.conv2.conv.bias),
                              residual=srcs[0], want_stats=not upsample_out, norm=n2, upsample_out=upsample_out)
        pw, bias = self._fused_tail_weight(split)
        return ops.conv2d([h] + srcs, pw, bias=bias, want_stats=not upsample_out,
                          norm=None if n2 is None else n2 + [None] * len(srcs), upsample_out=upsample_out)

    @staticmethod
    def _make_norm(norm_type: str, channels: int, norm_groups: int, norm_eps: float) -> nn.Module:
        kind = norm_type.lower()
        if kind == "gn":
            return make_group_norm(channels, groups=norm_groups, eps=norm_eps)
        if kind == "rmsnorm":
            return RMSNormND(channels)
        raise ValueError(f"Unsupported norm_type '{norm_type}'")

    @staticmethod
    def _make_act(act: str):
        kind = act.lower()
        if kind in ("silu", "swish"):
            return nn.SiLU()
        if kind == "relu":
            return nn.ReLU()
        if kind == "gelu":
            return nn.GELU()
        raise ValueError(f"Unsupported activation '{act}'")


def build_resblock_gn_silu(**kwargs) -> ResBlockND:
    return ResBlockND(norm_type="gn", act="silu", **kwargs)


def build_resblock_gn_swish(**kwargs) -> ResBlockND:
    return ResBlockND(norm_type="gn", act="swish", **kwargs)


def build_resblock_rmsnorm_silu(**kwargs) -> ResBlockND:
    return ResBlockND(norm_type="rmsnorm", act="silu", **kwargs)


def build_resblock_rmsnorm_swish(**kwargs) -> ResBlockND:
    return ResBlockND(norm_type="rmsnorm", act="swish", **kwargs)
