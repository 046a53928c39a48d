"""API-parity shells of `src/nn/ops/pooling.py` (patchify pooling / plain pooling).  None of them is reached by a
BASELINE config (out of scope): calling them raises OutOfScopeError."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from ..._runtime import out_of_scope
from .convolution import ConvND, ConvTransposeND

SizeArg = Union[int, Tuple[int, ...]]


def _is_unit(factor) -> bool:
    return factor == 1 or (isinstance(factor, (tuple, list)) and all(f == 1 for f in factor))


class PoolND(nn.Module):
    """Patchify: strided conv with kernel = stride = pool_factor (`.down`)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, pool_factor: SizeArg = 2):
        super().__init__()
        self.down = nn.Identity() if _is_unit(pool_factor) else ConvND(
            spatial_dims, in_channels, out_channels, kernel_size=pool_factor, stride=pool_factor, padding=0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down(x)


class UnPoolND(nn.Module):
    """Un-patchify: transposed conv with kernel = stride = pool_factor (`.up`)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, pool_factor: SizeArg = 2):
        super().__init__()
        self.up = nn.Identity() if _is_unit(pool_factor) else ConvTransposeND(
            spatial_dims, in_channels, out_channels, kernel_size=pool_factor, stride=pool_factor, padding=0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.up(x)


class AvgPoolND(nn.Module):
    def __init__(self, spatial_dims: int, kernel_size: SizeArg = 2, stride: Optional[SizeArg] = None,
                 padding: SizeArg = 0):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        self.pool = {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d}[spatial_dims](
            kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out_of_scope("AvgPoolND")


class MaxPoolND(nn.Module):
    def __init__(self, spatial_dims: int, kernel_size: SizeArg = 2, stride: Optional[SizeArg] = None,
                 padding: SizeArg = 0, dilation: SizeArg = 1, return_indices: bool = False, ceil_mode: bool = False):
        super().__init__()
        if spatial_dims not in (1, 2, 3):
            raise ValueError("spatial_dims must be 1, 2 or 3")
        self.pool = {1: nn.MaxPool1d, 2: nn.MaxPool2d, 3: nn.MaxPool3d}[spatial_dims](
            kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation,
            return_indices=return_indices, ceil_mode=ceil_mode)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out_of_scope("MaxPoolND")
