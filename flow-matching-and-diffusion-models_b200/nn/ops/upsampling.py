"""UpsampleND / DownsampleND (`src/nn/ops/upsampling.py:8-62`) on the B200 kernels (2-D, learned resampling)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import out_of_scope
from .convolution import ConvND


class UpsampleND(nn.Module):
    """Nearest x2 then 3x3 conv (`.conv`)."""

    def __init__(self, spatial_dims: int, channels: int, use_conv: bool = True):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        self.channels, self.use_conv, self.spatial_dims = channels, use_conv, spatial_dims
        if use_conv:
            self.conv = ConvND(spatial_dims, channels, channels, kernel_size=3, padding=1)

    def can_fuse_into_producer(self) -> bool:
        """True when the producer of the input may store it already upsampled (`ops.conv2d(upsample_out=True)`) and
        call `forward(x, upsampled=True)`."""
        return self.spatial_dims == 2 and self.channels % 8 == 0 and self.use_conv and self.conv.fast_path_ok()

    def forward(self, x: torch.Tensor, upsampled: bool = False) -> torch.Tensor:
        assert x.shape[1] == self.channels
        if upsampled:
            return self.conv(ops.to_nhwc_bf16(x), want_stats=True)
        if self.spatial_dims != 2 or self.channels % 8:
            out_of_scope(f"UpsampleND(spatial_dims={self.spatial_dims}, channels={self.channels})")
        y = ops.upsample_nearest2x(ops.to_nhwc_bf16(x))
        return self.conv(y, want_stats=True) if self.use_conv else y


class DownsampleND(nn.Module):
    """3x3 stride-2 pad-1 conv (`.op`) or, without conv, 2x average pooling (out of scope)."""

    def __init__(self, spatial_dims: int, channels: int, use_conv: bool = True):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        self.channels, self.use_conv, self.spatial_dims = channels, use_conv, spatial_dims
        if use_conv:
            self.op = ConvND(spatial_dims, in_channels=channels, out_channels=channels, kernel_size=3, stride=2,
                             padding=1)
        else:
            self.op = {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d}[spatial_dims](kernel_size=2, stride=2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[1] == self.channels
        if not self.use_conv:
            out_of_scope("DownsampleND(use_conv=False)")
        return self.op(x, want_stats=True)
