"""Shared test fixtures (GPU): seeded synthetic LDCT-shaped pairs and a briefly TRAINED epsilon denoiser.

Why train inside a test: DDIM / DPM-Solver++ divide the prediction by sqrt(abar_t) (~ 1/160 at t = 980) and clip, so a
random-init epsilon network turns the trajectory into clamped sign patterns and a final-sample PSNR says nothing about
the sampler.  A few hundred epsilon-target steps (`diffusion_lib.py:141-185`, through `fmdm_b200.training.
DiffusionTrainer`, seeded, deterministic kernels) give a conditioned denoiser on which the north-star bar (>= 40 dB
against the fp32 oracle run on the SAME trained weights) is meaningful.  The weights never leave the GPU box; the
recipe below is the committed, seeded script that regenerates them."""
import torch

DEV = "cuda"


def synthetic_pair(b: int, hw: int, gen: torch.Generator):
    """clean: smooth random field in [0,1] (bicubic-upsampled uniform noise); ldct: clean + 0.05 N(0,1), clamped -
    the synthetic conditioning of SURVEY.md 8d (low-dose = noisy version of the target)."""
    low = torch.rand(b, 1, max(2, hw // 16), max(2, hw // 16), generator=gen, device=DEV)
    clean = torch.nn.functional.interpolate(low, size=(hw, hw), mode="bicubic", align_corners=False).clamp_(0, 1)
    ldct = (clean + 0.05 * torch.randn(b, 1, hw, hw, generator=gen, device=DEV)).clamp_(0, 1)
    return clean, ldct


def train_epsilon_denoiser(model, *, hw: int, batch: int, steps: int, lr: float = 1e-4, warmup: int = 100,
                           seed: int = 123, beta_start: float = 1e-4, beta_end: float = 0.02):
    """AdamW(lr, linear warm-up) epsilon-target training on `synthetic_pair` batches; returns the last losses' mean.
    (the reference recipe: lr 1e-4, warm-up, AdamW, `configs/LDCT/LDCT_ddpm_diffusers_nd.json:12-15`)"""
    from fmdm_b200.pipelines.utils import build_scheduler
    from fmdm_b200.training import DiffusionTrainer

    sched, _ = build_scheduler({"name": "ddpm", "params": {"beta_start": beta_start, "beta_end": beta_end}}, {})
    tr = DiffusionTrainer(model, sched, lr=lr, weight_decay=0.0)
    gen = torch.Generator(device=DEV).manual_seed(seed)
    losses = []
    for i in range(steps):
        clean, ldct = synthetic_pair(batch, hw, gen)
        tr.optimizer.param_groups[0]["lr"] = lr * min(1.0, (i + 1) / max(1, warmup))
        losses.append(tr.step(clean, ldct))
    tr.reducer.remove()
    model.eval()
    tail = losses[-10:]
    return float(sum(float(l) for l in tail) / len(tail))
