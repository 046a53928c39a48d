"""One optimisation step of flow-matching training: the loop body of `src/pipelines/train/flow_matching_lib.py:138-182`
(noise / time sampling, x_t mixing, conditioning concat, denoiser forward, MSE on the velocity target, backward,
optimiser step), with the data-parallel gradient all-reduce of BASELINE config 5 overlapped with the backward."""
from __future__ import annotations

import math
from typing import Optional

import torch

from .. import ops
from . import functions as F
from .ddp import BucketedAllReduce
from .optim import FusedAdamW


def flow_matching_loss(model, clean: torch.Tensor, ldct: Optional[torch.Tensor], *, noise=None, t=None,
                       num_train_timesteps: int = 1000) -> torch.Tensor:
    """`flow_matching_lib.py:150-164`: x_t = (1-t) clean + t noise, target = noise - clean, loss = MSE(model(x_t), target).

    `noise` / `t` may be passed in (tests, seeded benchmarks); otherwise they are drawn as the reference draws them."""
    if noise is None:
        noise = torch.randn_like(clean)
    if t is None:
        t = torch.rand(clean.size(0), device=clean.device)
    timesteps = (t * (num_train_timesteps - 1)).long()
    x_t = ops.sched_add_noise(clean.float().contiguous(), noise.float().contiguous(), (1.0 - t).float().contiguous(),
                              t.float().contiguous())
    pred = model(x_t, timesteps, context=ldct) if ldct is not None else model(x_t, timesteps)
    return F.mse_loss(pred, noise, clean)


class FlowMatchingTrainer:
    """Owns the optimiser and the gradient reducer of one rank; `step()` is one `optimizer.step()` worth of work."""

    def __init__(self, model, *, lr: float = 1e-4, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_accum: int = 1, num_train_timesteps: int = 1000, bucket_bytes: int = 64 << 20, group=None):
        self.model = model
        self.optimizer = FusedAdamW(model.parameters(), lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
        self.reducer = BucketedAllReduce(self.optimizer.flat, bucket_bytes=bucket_bytes, group=group)
        self.optimizer.grad_scale = 1.0 / self.reducer.world
        self.grad_accum = max(1, int(grad_accum))
        self.num_train_timesteps = int(num_train_timesteps)

    def step(self, clean: torch.Tensor, ldct: Optional[torch.Tensor] = None, *, noise=None, t=None) -> torch.Tensor:
        """Returns the (detached, device-resident) mean loss of this rank's batch."""
        self.model.train()
        bs = clean.size(0)
        chunk = max(1, math.ceil(bs / self.grad_accum))
        cc = clean.split(chunk)
        lc = ldct.split(chunk) if ldct is not None else [None] * len(cc)
        nc = noise.split(chunk) if noise is not None else [None] * len(cc)
        tc = t.split(chunk) if t is not None else [None] * len(cc)
        self.optimizer.zero_grad()
        total = None
        for i, (c, l, n, tt) in enumerate(zip(cc, lc, nc, tc)):
            if i == len(cc) - 1:
                self.reducer.arm()
            loss = flow_matching_loss(self.model, c, l, noise=n, t=tt, num_train_timesteps=self.num_train_timesteps)
            (loss / self.grad_accum).backward()
            w = loss.detach() * (c.size(0) / bs)
            total = w if total is None else total + w
        self.reducer.finish()
        self.optimizer.step()
        return total
