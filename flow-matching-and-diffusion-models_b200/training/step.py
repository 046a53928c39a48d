"""One optimisation step of flow-matching training: the loop body of `src/pipelines/train/flow_matching_lib.py:138-182`
(noise / time sampling, x_t mixing, conditioning concat, denoiser forward, MSE on the velocity target, backward,
optimiser step), with the data-parallel gradient all-reduce of BASELINE config 5 overlapped with the backward - and
its epsilon-target twin, the loop body of `src/pipelines/train/diffusion_lib.py:141-185` (`DiffusionTrainer`)."""
from __future__ import annotations

import math
from typing import Optional

import torch

from .. import ops
from . import functions as F
from .ddp import BucketedAllReduce
from .optim import FusedAdamW


def _forward(model, x, timesteps, ldct, conditioning: str, latent_norm):
    """`flow_matching_lib.py:154-164` / `diffusion_lib.py:159-170`: "concatenate" hands the conditioning image to the
    stem (virtual concat); "attention" normalises the conditioning latents and hands them to the cross-attention blocks."""
    if ldct is None:
        return model(x, timesteps)
    if str(conditioning).lower() == "attention":
        from ..pipelines.utils import normalize_latent_conditioning

        return model(x, timesteps, context_ca=normalize_latent_conditioning(ldct, latent_norm))
    return model(x, timesteps, context=ldct)


def flow_matching_loss(model, clean: torch.Tensor, ldct: Optional[torch.Tensor], *, noise=None, t=None,
                       num_train_timesteps: int = 1000, conditioning: str = "concatenate",
                       latent_norm: Optional[str] = None) -> torch.Tensor:
    """`flow_matching_lib.py:150-164`: x_t = (1-t) clean + t noise, target = noise - clean, loss = MSE(model(x_t), target).

    `noise` / `t` may be passed in (tests, seeded benchmarks); otherwise they are drawn as the reference draws them."""
    if noise is None:
        noise = torch.randn_like(clean)
    if t is None:
        t = torch.rand(clean.size(0), device=clean.device)
    timesteps = (t * (num_train_timesteps - 1)).long()
    x_t = ops.sched_add_noise(clean.float().contiguous(), noise.float().contiguous(), (1.0 - t).float().contiguous(),
                              t.float().contiguous())
    pred = _forward(model, x_t, timesteps, ldct, conditioning, latent_norm)
    return F.mse_loss(pred, noise, clean)


def diffusion_loss(model, clean: torch.Tensor, ldct: Optional[torch.Tensor], sqrt_ac: torch.Tensor,
                   sqrt_1m_ac: torch.Tensor, *, noise=None, timesteps=None, conditioning: str = "concatenate",
                   latent_norm: Optional[str] = None) -> torch.Tensor:
    """`diffusion_lib.py:153-171`: timesteps ~ U{0..T-1}, noisy = scheduler.add_noise(clean, noise, timesteps)
    (= sqrt(abar_t) clean + sqrt(1-abar_t) noise), loss = MSE(model(noisy, timesteps), noise).

    `sqrt_ac` / `sqrt_1m_ac`: the scheduler's fp32 sqrt(abar) / sqrt(1-abar) tables on the device (gathered there, so
    the step has no host round trip and can be captured into the step graph)."""
    if noise is None:
        noise = torch.randn_like(clean)
    if timesteps is None:
        timesteps = torch.randint(0, sqrt_ac.numel(), (clean.size(0),), device=clean.device).long()
    noisy = ops.sched_add_noise(clean.float().contiguous(), noise.float().contiguous(),
                                sqrt_ac[timesteps].contiguous(), sqrt_1m_ac[timesteps].contiguous())
    pred = _forward(model, noisy, timesteps, ldct, conditioning, latent_norm)
    return F.mse_loss(pred, noise)


class FlowMatchingTrainer:
    """Owns the optimiser and the gradient reducer of one rank; `step()` is one `optimizer.step()` worth of work.

    cuda_graph=True (default): after `graph_warmup` eager steps on a fixed batch shape, zero_grad + noise/time sampling +
    forward + loss + backward are captured into ONE CUDA graph and replayed each step (the eager step issues ~2500
    small launches and is host-bound); the gradient all-reduce (data parallel) and the single optimiser kernel run
    after the replay.  Eager steps (shape changes, caller-supplied noise/t, gradient accumulation) keep the all-reduce
    overlapped with the backward through the bucket hooks."""

    def __init__(self, model, *, lr: float = 1e-4, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_accum: int = 1, num_train_timesteps: int = 1000, bucket_bytes: int = 64 << 20, group=None,
                 cuda_graph: bool = True, graph_warmup: int = 2, backward_cut="auto", side_streams: bool = True,
                 allreduce_max_ctas: Optional[int] = None, conditioning: str = "concatenate",
                 latent_norm: Optional[str] = None):
        import torch.distributed as dist

        from .graph import flat_param_order, supported

        import os

        self.model = model
        self.conditioning, self.latent_norm = str(conditioning), latent_norm    # `training.conditioning` / `.latent_norm`
        self.side_streams = bool(side_streams)
        # a communicator of its own for the gradient all-reduce, capped at a few CTAs: the collective runs beside the
        # second backward stage, where every SM it occupies is taken from the convs (NCCL's default sizes for speed)
        if allreduce_max_ctas is None and os.environ.get("FMDM_ALLREDUCE_MAX_CTAS"):
            allreduce_max_ctas = int(os.environ["FMDM_ALLREDUCE_MAX_CTAS"])
        if (allreduce_max_ctas and group is None and dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1 and dist.get_backend() == "nccl"):
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = int(allreduce_max_ctas)
            opts.config.min_ctas = 1
            group = dist.new_group(backend="nccl", pg_options=opts)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # two-stage backward (`training.graph.BackwardCut`): "auto" = cut when gradients are all-reduced, so the
        # reduction of the late layers' gradients overlaps the early layers' backward; an int forces the cut position
        if backward_cut == "auto":
            backward_cut = self._default_cut(model) if world > 1 else None
        self.backward_cut = backward_cut
        order = flat_param_order(model, backward_cut) if supported(model) else None
        self.optimizer = FusedAdamW(model.parameters(), lr=lr, weight_decay=weight_decay, betas=betas, eps=eps,
                                    flat_order=order)
        self.reducer = BucketedAllReduce(self.optimizer.flat, bucket_bytes=bucket_bytes, group=group)
        if self.reducer.world > 1:
            # what torch DDP does at construction: every replica starts from rank 0's parameters and buffers (ranks
            # that seeded differently, or where only rank 0 loaded a checkpoint, would otherwise diverge silently)
            import torch.distributed as dist

            dist.broadcast(self.optimizer.flat.data, src=dist.get_global_rank(group, 0) if group is not None else 0,
                           group=group)
            for buf in model.buffers():
                dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            torch._C._increment_version(self.optimizer.flat.params)
        F.ensure_tickets(self.optimizer.flat.data.device)
        self.optimizer.grad_scale = 1.0 / self.reducer.world
        self.grad_accum = max(1, int(grad_accum))
        self.num_train_timesteps = int(num_train_timesteps)
        self.cuda_graph = bool(cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self._graph = None
        self._graph2 = None
        self._graph_key = None
        self._eager_steps = 0
        self._static = None
        self._stage_ranges = None
        self.reduce_mode = "single rank" if self.reducer.world == 1 else "after the graph replay"

    # ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _default_cut(model):
        """Cut after the down blocks that run above 1/16 of the input resolution: in the LDCT denoisers that leaves
        ~5 % of the parameters (and ~40 % of the backward time) to the second stage."""
        if hasattr(model, "down_blocks"):
            n = len(model.down_blocks)
            return n - 2 if n >= 3 else None
        if hasattr(model, "input_blocks"):
            levels = len(getattr(model, "channel_mult", ()) or ())
            per_level = int(getattr(model, "num_res_blocks", 0)) + 1
            return 1 + per_level * (levels - 2) if levels >= 3 and per_level > 1 else None
        return None

    def _backward(self, loss) -> None:
        """`loss.backward()`; with a cut forward the two stages run back to back."""
        from .graph import finish_backward

        with self._direct_grads():
            loss.backward()
            finish_backward(self.model)
        F.assert_slots_drained()

    def _loss(self, clean, ldct, noise, t) -> torch.Tensor:
        return flow_matching_loss(self.model, clean, ldct, noise=noise, t=t,
                                  num_train_timesteps=self.num_train_timesteps, conditioning=self.conditioning,
                                  latent_norm=self.latent_norm)

    def _direct_grads(self):
        """Backward kernels write parameter gradients straight into the flat buffer (`functions.DIRECT_PARAM_GRADS`):
        valid when the optimiser step follows ONE backward (no gradient accumulation across micro-batches)."""
        trainer = self

        class _Ctx:
            def __enter__(self):
                self.saved = (F.DIRECT_PARAM_GRADS, F.GRAD_READY_HOOK, F.SIDE_STREAMS)
                F.DIRECT_PARAM_GRADS = trainer.grad_accum == 1
                F.GRAD_READY_HOOK = trainer.reducer._on_grad if trainer.grad_accum == 1 else None
                # parameter-gradient work of small layers on a side stream: only where no bucket all-reduce is launched
                # from the gradient-ready hooks in the middle of the backward (they would have to wait for it)
                F.SIDE_STREAMS = (trainer.side_streams and F.DIRECT_PARAM_GRADS
                                  and not (trainer.reducer._armed and trainer.reducer.world > 1))

            def __exit__(self, *exc):
                F.join_side_streams()
                F.DIRECT_PARAM_GRADS, F.GRAD_READY_HOOK, F.SIDE_STREAMS = self.saved
                return False

        return _Ctx()

    def _eager_step(self, clean, ldct, noise, t) -> torch.Tensor:
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
            with self._direct_grads():      # held over the forward too: `functions.fused_param` views
                loss = self._loss(c, l, n, tt)
            self._backward(loss / self.grad_accum)
            w = loss.detach() * (c.size(0) / bs)
            total = w if total is None else total + w
        self.reducer.finish()
        self.optimizer.step()
        return total

    def _capture(self, clean, ldct) -> None:
        """One graph for the whole step - or, with a backward cut, two graphs sharing a memory pool: [zero_grad,
        forward, loss, backward stage 1] and [backward stage 2]; the parameters whose gradients are complete after
        stage 1 are recorded during the capture (they become the all-reduce ranges launched between the replays)."""
        from .graph import finish_backward

        self._static = (clean.clone(), None if ldct is None else ldct.clone())
        sc, sl = self._static
        graph = torch.cuda.CUDAGraph()
        self.optimizer.flat.ensure_grad_views()
        cut = self.model.__dict__.get("_fm_backward_cut") is not None
        self.reducer.record(True)
        with torch.cuda.graph(graph):
            self.optimizer.zero_grad()
            with self._direct_grads():
                loss = self._loss(sc, sl, None, None)
                loss.backward()
            self._static_loss = loss.detach()
        early = self.reducer.record(False)
        self._graph, self._graph2, self._stage_ranges = graph, None, None
        if cut and "_fm_cut_state" in self.model.__dict__:
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2, pool=graph.pool()):
                with self._direct_grads():
                    finish_backward(self.model)
            self._graph2 = graph2
            first = self.reducer.ranges_of(early)
            self._stage_ranges = (first, self.reducer.complement(first))
            nb = [4 * sum(e - b for b, e in r) for r in self._stage_ranges]
            self.reduce_mode = (f"two-stage backward: {nb[0]} gradient bytes all-reduced while backward stage 2 runs, "
                                f"{nb[1]} bytes after it")
        F.assert_slots_drained()

    def step(self, clean: torch.Tensor, ldct: Optional[torch.Tensor] = None, *, noise=None, t=None) -> torch.Tensor:
        """Returns the (detached, device-resident) mean loss of this rank's batch."""
        self.model.train()
        self.model.__dict__["_fm_backward_cut"] = self.backward_cut
        try:
            return self._step(clean, ldct, noise, t)
        finally:
            self.model.__dict__.pop("_fm_backward_cut", None)

    def _step(self, clean, ldct, noise, t) -> torch.Tensor:
        key = (tuple(clean.shape), None if ldct is None else tuple(ldct.shape), clean.dtype)
        graphable = (self.cuda_graph and noise is None and t is None and self.grad_accum == 1 and clean.is_cuda)
        if not graphable or key != self._graph_key:
            if key != self._graph_key:
                self._graph, self._graph2, self._graph_key, self._eager_steps = None, None, key, 0
            if not graphable or self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._eager_step(clean, ldct, noise, t)
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._eager_step(clean, ldct, noise, t)
            self._capture(clean, ldct)
        sc, sl = self._static
        sc.copy_(clean, non_blocking=True)
        if sl is not None:
            sl.copy_(ldct, non_blocking=True)
        self._graph.replay()
        if self._graph2 is not None:
            first, rest = self._stage_ranges
            self.reducer.launch(first)      # overlaps the second backward stage
            self._graph2.replay()
            self.reducer.launch(rest)
            self.reducer.wait()
        elif self.reducer.world > 1:
            self.reducer.reduce_all()
        self.optimizer.step()
        return self._static_loss.clone()


class DiffusionTrainer(FlowMatchingTrainer):
    """The epsilon-target step of `src/pipelines/train/diffusion_lib.py:141-185`: same optimiser / reducer / step-graph
    machinery as `FlowMatchingTrainer`, the loss is `diffusion_loss` over the noise schedule of `scheduler` (any object
    with `alphas_cumprod` and `config.num_train_timesteps`: `DDPMScheduler`, `DDIMScheduler`, ...).  `step(clean, ldct,
    noise=..., t=...)`: `t` are the integer timesteps."""

    def __init__(self, model, scheduler, **kw):
        ac = getattr(scheduler, "alphas_cumprod", None)
        if ac is None:
            raise ValueError("DiffusionTrainer needs a scheduler with `alphas_cumprod` (ddpm / ddim / dpm_multistep)")
        kw.setdefault("num_train_timesteps", int(scheduler.config.num_train_timesteps))
        super().__init__(model, **kw)
        dev = next(model.parameters()).device
        ac = ac.to(torch.float32)
        self.sqrt_ac = (ac ** 0.5).to(dev).contiguous()
        self.sqrt_1m_ac = ((1 - ac) ** 0.5).to(dev).contiguous()

    def _loss(self, clean, ldct, noise, t) -> torch.Tensor:
        return diffusion_loss(self.model, clean, ldct, self.sqrt_ac, self.sqrt_1m_ac, noise=noise, timesteps=t,
                              conditioning=self.conditioning, latent_norm=self.latent_norm)
