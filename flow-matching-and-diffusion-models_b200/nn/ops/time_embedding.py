"""Sinusoidal timestep embedding (`src/nn/ops/time_embedding.py:4-32`) on the sm_100a kernel (accurate sin/cos)."""
from __future__ import annotations

import torch

from ... import ops


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000, *, flip_sin_to_cos: bool = True,
                       freq_shift: int = 0) -> torch.Tensor:
    """(N,) timesteps -> (N, dim) fp32 features laid out [sin | cos] (or [cos | sin] when flip_sin_to_cos)."""
    return ops.timestep_embedding(timesteps, dim, float(max_period), flip_sin_to_cos=flip_sin_to_cos,
                                  freq_shift=float(freq_shift))
