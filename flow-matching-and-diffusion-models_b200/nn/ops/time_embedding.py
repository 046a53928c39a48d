"""Sinusoidal timestep embedding (`src/nn/ops/time_embedding.py:4-32`) on the sm_100a kernel (accurate sin/cos)."""
from __future__ import annotations

import torch

from ... import ops


def timestep_embedding(timesteps, dim: int, max_period: int = 10000, *, flip_sin_to_cos: bool = True,
                       freq_shift: int = 0, batch=None, t_table=None, step_dev=None) -> torch.Tensor:
    """(N,) timesteps -> (N, dim) fp32 features laid out [sin | cos] (or [cos | sin] when flip_sin_to_cos).

    Graph-replay form: `timesteps=None, batch=N, t_table=<fp32 device table>, step_dev=<int32 device cursor>`."""
    return ops.timestep_embedding(timesteps, dim, float(max_period), flip_sin_to_cos=flip_sin_to_cos,
                                  freq_shift=float(freq_shift), batch=batch, t_table=t_table, step_dev=step_dev)
