"""ConvND — drop-in for `src/nn/ops/convolution.py:8-54` whose 2-D forward runs the tcgen05 implicit-GEMM kernel.

Same constructor, same child (`.conv`, an `nn.Conv2d` that only stores parameters => identical state_dict keys).
forward() accepts a logical NCHW tensor (fp32/bf16, any layout) or a tuple of tensors read as a virtual channel
concat, and returns a bf16 channels_last tensor of logical shape (B, Cout, Ho, Wo)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import ParamCache, f32, out_of_scope

SizeArg = Union[int, Tuple[int, ...]]


def _as_int(v) -> Optional[int]:
    if isinstance(v, int):
        return v
    v = tuple(v)
    return v[0] if all(e == v[0] for e in v) else None


class ConvND(nn.Module):
    weight_split = False  # split-bf16 weights (ops.PackedConvWeight); set per model by BaseUNetND.set_weight_split

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: SizeArg = 3,
                 stride: SizeArg = 1, padding: Optional[SizeArg] = None, dilation: SizeArg = 1, groups: int = 1,
                 bias: bool = True):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        if padding is None:
            padding = kernel_size // 2 if isinstance(kernel_size, int) else tuple(k // 2 for k in kernel_size)
        ctor = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
        self.conv = ctor(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)
        self.spatial_dims = spatial_dims
        self._cache = ParamCache()

    # ---- what the sm_100a kernel covers --------------------------------------------------------------------
    def fast_path_ok(self) -> bool:
        c = self.conv
        k, s, p, d = _as_int(c.kernel_size), _as_int(c.stride), _as_int(c.padding), _as_int(c.dilation)
        return (self.spatial_dims == 2 and k in (1, 3) and s in (1, 2) and p == (k // 2 if k else None) and d == 1
                and c.groups == 1 and c.in_channels % 8 == 0 and c.out_channels % 8 == 0
                and not (k == 1 and s != 1))

    def packed(self, split: Sequence[int]):
        """K-major bf16 weight with one K segment per entry of `split` (channel counts of the virtual concat)."""
        w = self.conv.weight

        def build():
            parts, c0 = [], 0
            for c in split:
                parts.append((w, c0, c))
                c0 += c
            return ops.pack_conv_weight(parts, split=self.weight_split)

        return self._cache.get(f"w{int(self.weight_split)}:" + ",".join(map(str, split)), [w], build)

    def forward(self, x, *, addvec=None, residual=None, want_stats: bool = False) -> torch.Tensor:
        srcs = list(x) if isinstance(x, (tuple, list)) else [x]
        if not self.fast_path_ok():
            c = self.conv
            if self.spatial_dims == 2 and _as_int(c.kernel_size) == 3 and _as_int(c.stride) == 1 \
                    and _as_int(c.padding) == 1 and c.groups == 1 and _as_int(c.dilation) == 1 \
                    and addvec is None and residual is None:
                if c.in_channels <= 8 and c.out_channels % 8 == 0 and len(srcs) <= 2:
                    xs = [s.float() for s in srcs]
                    return ops.conv_stem(xs[0], xs[1] if len(xs) == 2 else None, f32(c.weight), f32(c.bias))
                if c.out_channels <= 4 and c.in_channels % 8 == 0 and len(srcs) == 1:
                    return ops.conv_head(ops.to_nhwc_bf16(srcs[0]), f32(c.weight), f32(c.bias))
            out_of_scope(f"ConvND(spatial_dims={self.spatial_dims}, k={c.kernel_size}, s={c.stride}, "
                         f"groups={c.groups}, {c.in_channels}->{c.out_channels})")
        srcs = [ops.to_nhwc_bf16(s) for s in srcs]
        pw = self.packed([s.shape[1] for s in srcs])
        return ops.conv2d(srcs, pw, stride=_as_int(self.conv.stride), bias=f32(self.conv.bias), addvec=addvec,
                          residual=None if residual is None else ops.to_nhwc_bf16(residual), want_stats=want_stats)


class ConvTransposeND(nn.Module):
    """API-parity shell for `src/nn/ops/convolution.py:57-96`; not reached by any BASELINE config (out of scope)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: SizeArg = 2,
                 stride: SizeArg = 2, padding: SizeArg = 0, output_padding: Optional[SizeArg] = None, groups: int = 1,
                 bias: bool = True):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        ctor = {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}[spatial_dims]
        self.convT = ctor(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                          output_padding=output_padding or 0, groups=groups, bias=bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out_of_scope("ConvTransposeND")
