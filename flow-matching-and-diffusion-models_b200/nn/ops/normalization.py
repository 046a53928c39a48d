"""GroupNorm construction (`src/nn/ops/normalization.py:11-19`) plus the fused GroupNorm(+SiLU) entry point."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import f32, out_of_scope


def make_group_norm(channels: int, groups: int = 32, eps: float = 1e-5) -> nn.GroupNorm:
    """GroupNorm whose group count falls back to the largest divisor of `channels` not above `groups`."""
    g = max(1, min(groups, channels))
    while g > 1 and channels % g:
        g -= 1
    return nn.GroupNorm(g, channels, eps=eps)


def fused_group_norm(norm: nn.GroupNorm, srcs: Sequence[torch.Tensor], *, silu: bool,
                     scale_shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run `norm` (+SiLU, + (1+scale)*y+shift) over the virtual concat of `srcs` with the K2 kernel."""
    srcs = [ops.to_nhwc_bf16(s) for s in srcs]
    if len(srcs) > 2:
        srcs = [ops.to_nhwc_bf16(torch.cat(srcs, 1))]
    if any(s.shape[1] % 8 for s in srcs):
        out_of_scope(f"GroupNorm over {[s.shape[1] for s in srcs]} channels (needs multiples of 8)")
    return ops.group_norm(srcs, norm.num_groups, norm.eps, f32(norm.weight), f32(norm.bias), silu=silu,
                          scale_shift=scale_shift)


def fused_group_norm_table(norm: nn.GroupNorm, srcs: Sequence[torch.Tensor], *, silu: bool,
                           scale_shift: Optional[torch.Tensor] = None):
    """`norm` (+SiLU, + scale-shift) as the per-(sample, channel) affine table a consumer conv applies inside its
    operand path (`ops.conv2d(..., norm=...)`), or None when the sources do not qualify."""
    if not 1 <= len(srcs) <= 2 or any(s.shape[1] % 8 for s in srcs):
        return None
    return ops.group_norm_table(list(srcs), norm.num_groups, norm.eps, f32(norm.weight), f32(norm.bias), silu=silu,
                                scale_shift=scale_shift)


class RMSNormND(nn.Module):
    """API-parity shell for `src/nn/ops/normalization.py:22-34` (not on any BASELINE path; out of scope)."""

    def __init__(self, channels: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out_of_scope("RMSNormND")
