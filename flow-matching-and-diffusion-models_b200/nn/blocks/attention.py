"""Attention blocks — drop-ins for `src/nn/blocks/attention.py` on the B200 kernels.

In scope: `DiffusersAttentionND` self-attention (GN -> fused QKV 1x1 GEMM on tcgen05 ->
K3 softmax(QK^T)V -> out-proj GEMM with the residual in its epilogue) and `SpatialSelfAttention` including the
reference's raw-memory head split.  Cross-attention (`SpatialCrossAttention`, `DiffusersAttentionND`
with `context_dim`): queries as above, keys/values from `ops.context_kv` (GroupNorm over the context tokens + the tiny-K
projection in one pass, computed once per context tensor and reused over the sampling steps), K3 with Tq != Tk.
Linear attention keeps the API but is out of scope.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import ParamCache, f32, out_of_scope
from .common import zero_module


class QKVAttention(nn.Module):
    """softmax(q k^T / sqrt(d)) v over (N, heads, T, d) tensors (`attention.py:10-50`)."""

    def __init__(self, efficient_attn: bool = True, dropout: float = 0.0):
        super().__init__()
        self.dropout = dropout
        self.efficient_attn = bool(efficient_attn)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        if q.dim() != 4 or (self.training and self.dropout > 0) or q.shape[-1] not in (8, 16, 32, 64) \
                or v.shape[-1] != q.shape[-1]:
            out_of_scope(f"QKVAttention on {tuple(q.shape)} (dropout={self.dropout})")
        ops.require_cuda(q, "QKVAttention")
        b, h, tq, d = q.shape
        tk = k.shape[2]
        q, k, v = [t.to(torch.bfloat16).contiguous() for t in (q, k, v)]
        out = torch.empty((b, h, tq, d), dtype=torch.bfloat16, device=q.device)
        ops.attention(q, k, v, out, batch=b, heads=h, tq=tq, tk=tk, head_dim=d, q_strides=(h * tq * d, tq * d, d),
                      kv_strides=(h * tk * d, tk * d, d), o_strides=(h * tq * d, tq * d, d))
        return out


class LinearQKVAttention(nn.Module):
    """Softmax-factorised linear attention (`attention.py:53-70`), EfficientUNetND's default inside its levels
    (SURVEY §8f N4): one small CUDA-core kernel per call (O(T d^2))."""

    def __init__(self, dropout: float = 0.0, eps: float = 1e-6):
        super().__init__()
        self.dropout = dropout
        self.eps = eps

    def forward(self, q, k, v):
        if q.dim() != 4 or (self.training and self.dropout > 0) or q.shape[-1] not in (8, 16, 32, 64) \
                or v.shape[-1] != q.shape[-1] or not q.is_cuda:
            out_of_scope(f"LinearQKVAttention on {tuple(q.shape)} (dropout={self.dropout})")
        b, h, tq, d = q.shape
        tk = k.shape[2]
        q, k, v = [t.to(torch.bfloat16).contiguous() for t in (q, k, v)]
        out = torch.empty((b, h, tq, d), dtype=torch.bfloat16, device=q.device)
        ops.linear_attention(q, k, v, out, batch=b, heads=h, tq=tq, tk=tk, head_dim=d,
                             q_strides=(h * tq * d, tq * d, d), kv_strides=(h * tk * d, tk * d, d),
                             o_strides=(h * tq * d, tq * d, d), eps=self.eps)
        return out


def context_tokens(context: torch.Tensor, context_dim: int) -> torch.Tensor:
    """The context as (B, context_dim, T_ctx), with the reference's shape rules (`attention.py:160-175, 237-252`)."""
    if context.dim() == 3:
        if context.shape[1] == context_dim:
            return context
        if context.shape[-1] == context_dim:
            return context.transpose(1, 2)
        raise ValueError(f"Context channels mismatch: expected {context_dim}, got {tuple(context.shape)}.")
    if context.shape[1] != context_dim:
        raise ValueError(f"Context channels mismatch: expected {context_dim}, got {tuple(context.shape)}.")
    return context.reshape(context.shape[0], context.shape[1], -1)


class _ContextKV:
    """Keys/values of one context tensor: they depend on the context and the projection weights only, not on the
    sampling step, so they are computed once per (context object, parameter version) and reused."""

    def __init__(self):
        self._ref, self._sig, self._kv = None, None, None

    def get(self, context: torch.Tensor, params, build):
        sig = (context._version, context.data_ptr(), tuple(context.shape), ParamCache._sig(params))
        if self._ref is not None and self._ref() is context and self._sig == sig:
            return self._kv
        kv = build()
        import weakref

        if self._kv is not None and self._kv.shape == kv.shape and self._kv.device == kv.device:
            # refresh in place: a CUDA graph captured with the previous keys/values keeps reading this buffer
            self._kv.copy_(kv)
            kv = self._kv
        self._ref, self._sig, self._kv = weakref.ref(context), sig, kv
        return kv


CONTEXT_DIM_MAX = 16  # fm_context_kv_bf16 keeps a token's context vector / a weight row in registers


class ContextBlock(nn.Module):
    """Base class of layers consuming an external context tensor."""

    def forward(self, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise NotImplementedError


class SpatialSelfAttention(nn.Module):
    """CompVis-style MHSA block (`attention.py:82-117`): GN -> Conv1d qkv -> *raw reshape* head split -> SDPA ->
    raw reshape back -> zero-init Conv1d -> + x.  The raw reshape re-reads the channel-major (b, 3*inner, T) buffer
    as (b, heads, T, 3*dh); we reproduce it exactly by transposing the NHWC GEMM output to channel-major once and
    handing K3 the matching strides."""

    weight_split = False  # split-bf16 projection weights (see ConvND.weight_split)

    def __init__(self, dim: int, heads: int = 4, dim_head: int = 64, use_linear: bool = False,
                 use_efficient_attn: bool = True):
        super().__init__()
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.inner_dim = dim_head * heads
        self.use_linear = use_linear
        self.norm = nn.GroupNorm(max(1, math.gcd(dim, 32)), dim)
        self.qkv = nn.Conv1d(dim, self.inner_dim * 3, 1)
        self.attention = LinearQKVAttention() if use_linear else QKVAttention(efficient_attn=use_efficient_attn)
        self.proj_out = zero_module(nn.Conv1d(self.inner_dim, self.dim, 1))
        self._cache = ParamCache()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, c, *spatial = x.shape
        if len(spatial) != 2 or c % 8 or self.dim_head not in (8, 16, 32, 64):
            out_of_scope(f"SpatialSelfAttention(spatial={spatial}, dim_head={self.dim_head})")
        x = ops.to_nhwc_bf16(x)
        hh, ww = spatial
        t = hh * ww
        inner, dh, nh = self.inner_dim, self.dim_head, self.heads
        n = ops.group_norm([x], self.norm.num_groups, self.norm.eps, f32(self.norm.weight), f32(self.norm.bias),
                           silu=False)
        wq = self._cache.get(f"qkv{int(self.weight_split)}", [self.qkv.weight],
                             lambda: ops.pack_conv_weight([(self.qkv.weight.squeeze(-1), 0, c)],
                                                          split=self.weight_split))
        wo = self._cache.get(f"out{int(self.weight_split)}", [self.proj_out.weight],
                             lambda: ops.pack_conv_weight([(self.proj_out.weight.squeeze(-1), 0, inner)],
                                                          split=self.weight_split))
        qkv = ops.conv2d([n], wq, bias=f32(self.qkv.bias))                      # NHWC == [b][T][3*inner]
        qkv_cm = ops.transpose_bf16(qkv.permute(0, 2, 3, 1).reshape(b, t, 3 * inner))  # [b][3*inner][T]
        att = torch.empty((b, nh, t, dh), dtype=torch.bfloat16, device=x.device)
        flat = qkv_cm.view(-1)
        kernel = ops.linear_attention if self.use_linear else ops.attention
        kernel(flat, flat[dh:], flat[2 * dh:], att, batch=b, heads=nh, tq=t, tk=t, head_dim=dh,
               q_strides=(3 * inner * t, t * 3 * dh, 3 * dh), kv_strides=(3 * inner * t, t * 3 * dh, 3 * dh),
               o_strides=(nh * t * dh, t * dh, dh))
        # raw reshape (b, heads, T, dh) -> (b, inner, T), then back to NHWC for the projection GEMM
        h_tc = ops.transpose_bf16(att.view(b, inner, t))                          # [b][T][inner]
        h_nhwc = h_tc.view(b, hh, ww, inner).permute(0, 3, 1, 2)
        return ops.conv2d([h_nhwc], wo, bias=f32(self.proj_out.bias), residual=x, want_stats=True)

class SpatialCrossAttention(ContextBlock):
    """CompVis-style cross-attention block (`attention.py:120-189`, conditioning:"attention" configs):
    GN -> Conv1d q -> raw-reshape head split; context GN -> Conv1d kv (one small kernel, cached per context) -> raw
    reshape -> SDPA (Tq != Tk) -> raw reshape back -> zero-init Conv1d -> + x."""

    weight_split = False  # split-bf16 projection weights (see ConvND.weight_split)

    def __init__(self, dim: int, context_dim: int, heads: int = 4, dim_head: int = 64, use_linear: bool = False,
                 use_efficient_attn: bool = True):
        super().__init__()
        self.dim, self.context_dim, self.heads, self.dim_head = dim, context_dim, heads, dim_head
        self.inner_dim = dim_head * heads
        self.norm = nn.GroupNorm(max(1, math.gcd(dim, 32)), dim)
        self.context_norm = nn.GroupNorm(max(1, math.gcd(context_dim, 32)), context_dim)
        self.q_proj = nn.Conv1d(dim, self.inner_dim, 1)
        self.kv_proj = nn.Conv1d(context_dim, self.inner_dim * 2, 1)
        self.attention = LinearQKVAttention() if use_linear else QKVAttention(efficient_attn=use_efficient_attn)
        self.proj_out = zero_module(nn.Conv1d(self.inner_dim, self.dim, 1))
        self.use_linear = use_linear
        self._cache = ParamCache()
        self._kv = _ContextKV()

    def forward(self, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        if context is None:
            raise ValueError("SpatialCrossAttention requires a non-empty context tensor.")
        b, c, *spatial = x.shape
        if (len(spatial) != 2 or c % 8 or self.dim_head not in (8, 16, 32, 64)
                or self.context_dim > CONTEXT_DIM_MAX or not x.is_cuda):
            out_of_scope(f"SpatialCrossAttention(spatial={spatial}, dim_head={self.dim_head}, "
                         f"context_dim={self.context_dim})")
        x = ops.to_nhwc_bf16(x)
        hh, ww = spatial
        t, inner, dh, nh = hh * ww, self.inner_dim, self.dim_head, self.heads
        n = ops.group_norm([x], self.norm.num_groups, self.norm.eps, f32(self.norm.weight), f32(self.norm.bias),
                           silu=False)
        wq = self._cache.get(f"q{int(self.weight_split)}", [self.q_proj.weight],
                             lambda: ops.pack_conv_weight([(self.q_proj.weight.squeeze(-1), 0, c)],
                                                          split=self.weight_split))
        wo = self._cache.get(f"out{int(self.weight_split)}", [self.proj_out.weight],
                             lambda: ops.pack_conv_weight([(self.proj_out.weight.squeeze(-1), 0, inner)],
                                                          split=self.weight_split))
        q = ops.conv2d([n], wq, bias=f32(self.q_proj.bias))                            # NHWC == [b][T][inner]
        q_cm = ops.transpose_bf16(q.permute(0, 2, 3, 1).reshape(b, t, inner))          # [b][inner][T]
        kv_cm = self.precompute_context(context)
        tc = kv_cm.shape[-1]
        # the reference's raw reshapes: q (b, inner, T) -> (b, heads, T, dh); kv (b, 2*inner, Tc) -> (b, heads, Tc, 2*dh)
        att = torch.empty((b, nh, t, dh), dtype=torch.bfloat16, device=x.device)
        kvf = kv_cm.view(-1)
        kernel = ops.linear_attention if self.use_linear else ops.attention
        kernel(q_cm.view(-1), kvf, kvf[dh:], att, batch=b, heads=nh, tq=t, tk=tc, head_dim=dh,
               q_strides=(inner * t, t * dh, dh), kv_strides=(2 * inner * tc, tc * 2 * dh, 2 * dh),
               o_strides=(nh * t * dh, t * dh, dh))
        h_tc = ops.transpose_bf16(att.view(b, inner, t))                                # [b][T][inner]
        h_nhwc = h_tc.view(b, hh, ww, inner).permute(0, 3, 1, 2)
        return ops.conv2d([h_nhwc], wo, bias=f32(self.proj_out.bias), residual=x, want_stats=True)

    def precompute_context(self, context: torch.Tensor) -> torch.Tensor:
        """Keys/values of `context`, channel-major (b, 2*inner, Tc); cached per context object / parameter version."""
        cn = self.context_norm
        cf = context_tokens(context, self.context_dim)
        return self._kv.get(context, [self.kv_proj.weight, self.kv_proj.bias, cn.weight, cn.bias],
                            lambda: ops.context_kv(cf, f32(cn.weight), f32(cn.bias),
                                                   f32(self.kv_proj.weight.squeeze(-1)), f32(self.kv_proj.bias),
                                                   groups=cn.num_groups, eps=cn.eps, channel_major=True))

class DiffusersAttentionND(nn.Module):
    """Diffusers-style attention over flattened spatial tokens (`attention.py:192-274`), self-attention on the
    B200 kernels: q/k/v Linear layers run as ONE 1x1 implicit GEMM (N = 3C), to_out as another with the residual
    add in its epilogue.  Children / state_dict keys: group_norm, to_q, to_k, to_v, to_out.0."""

    weight_split = False  # split-bf16 projection weights (see ConvND.weight_split)

    def __init__(self, channels: int, heads: int = 1, context_dim: int | None = None, norm_num_groups: int = 32,
                 eps: float = 1e-5, dropout: float = 0.0, use_efficient_attn: bool = True):
        super().__init__()
        self.channels = channels
        self.heads = max(1, heads)
        self.head_dim = channels // self.heads
        self.context_dim = int(context_dim) if context_dim is not None else None
        self.group_norm = nn.GroupNorm(max(1, math.gcd(channels, norm_num_groups)), channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        kv_in = channels if self.context_dim is None else self.context_dim
        self.context_norm = None if self.context_dim is None else nn.GroupNorm(
            max(1, math.gcd(self.context_dim, norm_num_groups)), self.context_dim, eps=eps)
        self.to_k = nn.Linear(kv_in, channels)
        self.to_v = nn.Linear(kv_in, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(dropout)])
        self.attention = QKVAttention(efficient_attn=use_efficient_attn, dropout=dropout)
        self.dropout = dropout
        self._cache = ParamCache()
        self._kv = _ContextKV()

    def forward(self, hidden_states: torch.Tensor, context: torch.Tensor | None = None,
                upsample_out: bool = False) -> torch.Tensor:
        """upsample_out (fast path only): return the block output nearest-2x upsampled, folded into the store of the
        output projection (a following UpsampleND's F.interpolate)."""
        b, c = hidden_states.shape[:2]
        spatial = hidden_states.shape[2:]
        if self.context_dim is not None:
            if context is None:
                raise ValueError("DiffusersAttentionND cross-attention requires a non-empty context tensor.")
            if (len(spatial) != 2 or c % 8 or self.head_dim not in (8, 16, 32, 64) or self.context_dim > CONTEXT_DIM_MAX
                    or (self.training and self.dropout > 0) or not hidden_states.is_cuda):
                out_of_scope(f"DiffusersAttentionND cross-attention(spatial={tuple(spatial)}, head_dim={self.head_dim}, "
                             f"context_dim={self.context_dim})")
            return self._cross(hidden_states, context, upsample_out)
        if len(spatial) != 2 or c % 8 or self.head_dim not in (8, 16, 32, 64) or (self.training and self.dropout > 0):
            out_of_scope(f"DiffusersAttentionND(spatial={tuple(spatial)}, head_dim={self.head_dim})")
        x = ops.to_nhwc_bf16(hidden_states)
        hh, ww = spatial
        t = hh * ww
        gn = self.group_norm
        n = ops.group_norm([x], gn.num_groups, gn.eps, f32(gn.weight), f32(gn.bias), silu=False)

        def build_qkv():
            w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0).detach()
            bias = torch.cat([self.to_q.bias, self.to_k.bias, self.to_v.bias], 0).detach().float().contiguous()
            return ops.pack_conv_weight([(w, 0, c)], split=self.weight_split), bias

        wqkv, bqkv = self._cache.get(f"qkv{int(self.weight_split)}",
                                     [self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_q.bias,
                                             self.to_k.bias, self.to_v.bias], build_qkv)
        wout = self._cache.get(f"out{int(self.weight_split)}", [self.to_out[0].weight],
                               lambda: ops.pack_conv_weight([(self.to_out[0].weight, 0, c)], split=self.weight_split))
        qkv = ops.conv2d([n], wqkv, bias=bqkv)                                   # NHWC == [b][T][3C]
        flat = qkv.permute(0, 2, 3, 1).reshape(-1)
        att = ops.empty_nhwc(b, c, hh, ww, x.device)                             # [b][T][C]
        hd = self.head_dim
        ops.attention(flat, flat[c:], flat[2 * c:], att.permute(0, 2, 3, 1).reshape(-1), batch=b, heads=self.heads,
                      tq=t, tk=t, head_dim=hd, q_strides=(t * 3 * c, hd, 3 * c), kv_strides=(t * 3 * c, hd, 3 * c),
                      o_strides=(t * c, hd, c))
        return ops.conv2d([att], wout, bias=f32(self.to_out[0].bias), residual=x, want_stats=not upsample_out,
                          upsample_out=upsample_out)

    def _cross(self, hidden_states: torch.Tensor, context: torch.Tensor, upsample_out: bool) -> torch.Tensor:
        """`attention.py:232-274` with `context_dim`: q from the image tokens, k/v from the normalised context."""
        b, c, hh, ww = hidden_states.shape
        x = ops.to_nhwc_bf16(hidden_states)
        t, hd = hh * ww, self.head_dim
        gn = self.group_norm
        n = ops.group_norm([x], gn.num_groups, gn.eps, f32(gn.weight), f32(gn.bias), silu=False)
        wq = self._cache.get(f"q{int(self.weight_split)}", [self.to_q.weight],
                             lambda: ops.pack_conv_weight([(self.to_q.weight, 0, c)], split=self.weight_split))
        wout = self._cache.get(f"out{int(self.weight_split)}", [self.to_out[0].weight],
                               lambda: ops.pack_conv_weight([(self.to_out[0].weight, 0, c)], split=self.weight_split))
        q = ops.conv2d([n], wq, bias=f32(self.to_q.bias))                                # NHWC == [b][T][C]
        kv = self.precompute_context(context)                                            # [b][Tc][2C]: K | V
        tc = kv.shape[1]
        att = ops.empty_nhwc(b, c, hh, ww, x.device)
        kvf = kv.view(-1)
        ops.attention(q.permute(0, 2, 3, 1).reshape(-1), kvf, kvf[c:], att.permute(0, 2, 3, 1).reshape(-1), batch=b,
                      heads=self.heads, tq=t, tk=tc, head_dim=hd, q_strides=(t * c, hd, c),
                      kv_strides=(tc * 2 * c, hd, 2 * c), o_strides=(t * c, hd, c))
        return ops.conv2d([att], wout, bias=f32(self.to_out[0].bias), residual=x, want_stats=not upsample_out,
                          upsample_out=upsample_out)

    def precompute_context(self, context: torch.Tensor) -> torch.Tensor:
        """Keys/values of `context`, token-major (b, Tc, 2C); cached per context object / parameter version."""
        cn = self.context_norm
        cf = context_tokens(context, self.context_dim)

        def build_kv():
            w = torch.cat([self.to_k.weight, self.to_v.weight], 0).detach().float().contiguous()
            bias = torch.cat([self.to_k.bias, self.to_v.bias], 0).detach().float().contiguous()
            return ops.context_kv(cf, f32(cn.weight), f32(cn.bias), w, bias, groups=cn.num_groups, eps=cn.eps,
                                  channel_major=False)

        return self._kv.get(context, [self.to_k.weight, self.to_k.bias, self.to_v.weight, self.to_v.bias, cn.weight,
                                      cn.bias], build_kv)
