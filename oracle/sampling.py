"""ORACLE (test infrastructure, never imported by the product path).

CPU/GPU-agnostic restatement of the reference sampling loop `sample_with_scheduler`
(`/root/reference/src/pipelines/utils.py:163-220`) and of the scheduler selection done by `decode_diffusion_batch`
(`/root/reference/src/utils/model_utils/diffusion_utils.py:196-227`), over the oracle schedulers and any callable
denoiser `model(model_input, timesteps) -> prediction` (the reference's own modules or `oracle.denoiser`).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .schedulers import ORACLE_REGISTRY

_ALIASES = {  # pipelines/utils.py:74-84
    "flowmatch": ("flow_match_euler", {}),
    "flow_match_euler": ("flow_match_euler", {}),
    "ddim": ("ddim", {}),
    "ddpm": ("ddpm", {}),
    "dpmsolver1": ("dpm_multistep", {"solver_order": 1, "algorithm_type": "dpmsolver"}),
    "dpmsolver2": ("dpm_multistep", {"solver_order": 2, "algorithm_type": "dpmsolver"}),
    "dpmsolver++": ("dpm_multistep", {"solver_order": 2, "algorithm_type": "dpmsolver++"}),
    "unipc": ("unipc", {}),
}


def make_scheduler(name: str, num_train_timesteps: int = 1000, params: Optional[dict] = None):
    key, extra = _ALIASES.get(name, (name, {}))
    cls = ORACLE_REGISTRY[key]
    import inspect

    allowed = set(inspect.signature(cls.__init__).parameters) - {"self"}
    kw = {k: v for k, v in {**(params or {}), **extra}.items() if k in allowed}
    return cls(num_train_timesteps=num_train_timesteps, **kw)


def select_timesteps(timesteps: torch.Tensor, start_step=None, last_n_steps=None) -> torch.Tensor:
    if start_step is not None:
        if int(start_step) < 0:
            raise ValueError("start_step must be >= 0.")
        timesteps = timesteps[timesteps <= int(start_step)]
    if last_n_steps is not None:
        if int(last_n_steps) <= 0:
            raise ValueError("last_n_steps must be > 0.")
        timesteps = timesteps[-int(last_n_steps):]
    if timesteps.numel() == 0:
        raise ValueError("No timesteps selected after applying start_step/last_n_steps.")
    return timesteps


def sample_loop(model: Callable, scheduler, num_inference_steps: int, init_sample: torch.Tensor,
                cond: Optional[torch.Tensor] = None, start_step=None, last_n_steps=None,
                trace: Optional[list] = None) -> torch.Tensor:
    """for t in timesteps: pred = model(cat(x, cond), t.expand(B)); x = scheduler.step(pred, t, x).prev_sample"""
    scheduler.set_timesteps(num_inference_steps)
    timesteps = select_timesteps(scheduler.timesteps, start_step, last_n_steps)
    x = init_sample
    for t in timesteps:
        inp = x if cond is None else torch.cat([x, cond], dim=1)
        tt = t.to(x.device)
        if tt.dim() == 0:
            tt = tt.expand(x.size(0))
        pred = model(inp, tt)
        if trace is not None:
            trace.append(pred)
        x = scheduler.step(pred.to(x.device), t, x).prev_sample
    return x
