"""ORACLE (test infrastructure, not product code): CPU/fp32 restatement of one flow-matching optimisation step (and of
the epsilon-target diffusion step, `diffusion_loss`).

Follows `src/pipelines/train/flow_matching_lib.py` of the reference:
  :150-153  timesteps = (t * (num_train_timesteps - 1)).long();  x_t = (1 - t) * clean + t * noise
  :154-157  conditioning "concatenate": model_input = cat([x_t, ldct], dim=1)
  :162-164  pred = model(model_input, timesteps); target = noise - clean; loss = F.mse_loss(pred, target)
  :169,176  (loss / grad_accum).backward(); optimizer.step()  with  AdamW(model.parameters(), lr, weight_decay) (:74)
The denoiser is the functional restatement in oracle/denoiser.py (itself pinned to the reference's modules), driven
through torch autograd; AdamW is restated from torch.optim.AdamW's documented single-tensor update.

Pinned by tests/golden/train_step_*.pt, produced by oracle/make_golden_train.py from the reference's own model class,
torch autograd and torch.optim.AdamW (tests/test_oracle_training.py).  Only tests/, __graft_entry__.smoke() and the
bench's CPU legs may import this module."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import denoiser as OD

SD = Dict[str, torch.Tensor]


def flow_matching_loss(params: SD, cfg: dict, clean: torch.Tensor, ldct: Optional[torch.Tensor],
                       noise: torch.Tensor, t: torch.Tensor, num_train_timesteps: int = 1000):
    """flow_matching_lib.py:150-164 -> (loss, pred)."""
    timesteps = (t * (num_train_timesteps - 1)).long()
    tt = t[:, None, None, None]
    x_t = (1.0 - tt) * clean + tt * noise
    pred = OD.denoiser_forward(params, cfg, x_t, timesteps, conditioning="concatenate" if ldct is not None else None,
                               channels=clean.shape[1], context=ldct)
    return F.mse_loss(pred, noise - clean), pred


def diffusion_loss(params: SD, cfg: dict, clean: torch.Tensor, ldct: Optional[torch.Tensor], noise: torch.Tensor,
                   timesteps: torch.Tensor, alphas_cumprod: torch.Tensor):
    """The epsilon-target step of `src/pipelines/train/diffusion_lib.py:153-171` -> (loss, pred):
      :155-157  timesteps = randint(0, num_train_timesteps);  :158 noisy = scheduler.add_noise(clean, noise, timesteps)
                (diffusers DDPM/DDIM add_noise: sqrt(abar_t) * clean + sqrt(1 - abar_t) * noise)
      :161-162  conditioning "concatenate": model_input = cat([noisy, ldct], dim=1)
      :168-171  pred = model(model_input, timesteps);  loss = F.mse_loss(pred, noise)"""
    ac = alphas_cumprod.to(clean.device, torch.float32)
    sa = (ac[timesteps] ** 0.5)[:, None, None, None]
    sb = ((1 - ac[timesteps]) ** 0.5)[:, None, None, None]
    noisy = sa * clean + sb * noise
    pred = OD.denoiser_forward(params, cfg, noisy, timesteps, conditioning="concatenate" if ldct is not None else None,
                               channels=clean.shape[1], context=ldct)
    return F.mse_loss(pred, noise), pred


def loss_and_grads(sd: SD, cfg: dict, clean, ldct, noise, t, num_train_timesteps: int = 1000,
                   alphas_cumprod: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, SD]:
    """`alphas_cumprod` given: the epsilon-target diffusion step with integer timesteps `t`; else flow matching."""
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    if alphas_cumprod is not None:
        loss, _ = diffusion_loss(params, cfg, clean, ldct, noise, t, alphas_cumprod)
    else:
        loss, _ = flow_matching_loss(params, cfg, clean, ldct, noise, t, num_train_timesteps)
    loss.backward()
    return loss.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, *, lr: float,
               betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
    """torch.optim.AdamW (amsgrad=False, maximize=False): returns the updated (p, m, v)."""
    b1, b2 = betas
    p = p * (1.0 - lr * weight_decay)
    m = m + (1.0 - b1) * (g - m)
    v = v * b2 + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def train_steps(sd: SD, cfg: dict, batches, *, lr: float, weight_decay: float, num_train_timesteps: int = 1000):
    """Run len(batches) optimisation steps; returns (losses, final state_dict)."""
    sd = {k: v.detach().clone() for k, v in sd.items()}
    ms = {k: torch.zeros_like(v) for k, v in sd.items()}
    vs = {k: torch.zeros_like(v) for k, v in sd.items()}
    losses = []
    for step, (clean, ldct, noise, t) in enumerate(batches, 1):
        loss, grads = loss_and_grads(sd, cfg, clean, ldct, noise, t, num_train_timesteps)
        losses.append(float(loss))
        for k in sd:
            if sd[k].is_floating_point():
                sd[k], ms[k], vs[k] = adamw_step(sd[k], grads[k], ms[k], vs[k], step, lr=lr, weight_decay=weight_decay)
    return losses, sd
