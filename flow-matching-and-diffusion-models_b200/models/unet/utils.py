"""Timestep-feature helpers (`src/models/unet/utils.py`): the 2-layer time MLP runs as two tiny fp32 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ..._runtime import f32
from ...nn.ops.time_embedding import timestep_embedding


def time_mlp(x: torch.Tensor, lin1: nn.Linear, lin2: nn.Linear) -> torch.Tensor:
    """Linear -> SiLU -> Linear in fp32 (SiLU fused into the first kernel's epilogue)."""
    h = ops.linear_f32(x, f32(lin1.weight), f32(lin1.bias), silu_out=True)
    return ops.linear_f32(h, f32(lin2.weight), f32(lin2.bias))


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, out_channels)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(out_channels, out_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return time_mlp(x, self.linear_1, self.linear_2)


def build_timestep_features(timesteps, channels: int, *, max_period: int = 10000, flip_sin_to_cos: bool = True,
                            freq_shift: int = 0, batch=None, t_table=None, step_dev=None) -> torch.Tensor:
    return timestep_embedding(timesteps, channels, max_period=max_period, flip_sin_to_cos=flip_sin_to_cos,
                              freq_shift=freq_shift, batch=batch, t_table=t_table, step_dev=step_dev)
