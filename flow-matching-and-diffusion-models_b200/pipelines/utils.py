"""Sampling-loop helpers — the B200 counterpart of `src/pipelines/utils.py`.

Same public names and argument meaning as the reference (`SCHEDULER_REGISTRY`, `build_scheduler`,
`resolve_scheduler_override`, `sample_with_scheduler`, ...).  `sample_with_scheduler` keeps the reference's
semantics (`pipelines/utils.py:163-220`) but, for the in-scope models/schedulers, executes the N-step loop as a
replayed CUDA graph of the hand-written kernels: one capture of [denoiser forward -> fused scheduler step -> cursor
increment], N replays, no per-step host synchronisation (the reference syncs twice per step, `:211,214`).
"""
from __future__ import annotations

import inspect
import math
import time
from typing import Dict, Optional, Tuple

import torch

from .. import ops
from ..models.unet.base import BaseUNetND
from .schedulers import (DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler,
                         FlowMatchEulerDiscreteScheduler, UniPCMultistepScheduler, _SchedulerBase)

SCHEDULER_REGISTRY: Dict[str, type] = {
    "ddpm": DDPMScheduler,
    "ddim": DDIMScheduler,
    "dpm_multistep": DPMSolverMultistepScheduler,
    "unipc": UniPCMultistepScheduler,
    "flow_match_euler": FlowMatchEulerDiscreteScheduler,
    "flowmatch": FlowMatchEulerDiscreteScheduler,
}
# names the reference registers but the north star does not ask for
_OUT_OF_SCOPE_SCHEDULERS = ("dpm_sde",)


def resolve_conditioning_mode(value) -> Optional[str]:
    if value is None:
        return None
    text = str(value).strip().lower()
    return text or None


def build_scheduler(spec: Dict, training_cfg: Dict) -> Tuple[object, int]:
    """(scheduler, num_inference_steps) from the `model.scheduler` and `training` config dicts."""
    spec = dict(spec or {})
    training_cfg = dict(training_cfg or {})
    name = spec.get("name") or training_cfg.get("scheduler") or "ddpm"
    key = str(name).lower()
    if key in _OUT_OF_SCOPE_SCHEDULERS:
        raise NotImplementedError(
            f"fmdm_b200: scheduler '{name}' is outside the B200 hot path (built: flowmatch, ddpm, ddim, dpmsolver++, "
            "dpmsolver1/2, unipc)")
    if key not in SCHEDULER_REGISTRY:
        raise ValueError(f"Unknown scheduler '{name}'. Available: {', '.join(SCHEDULER_REGISTRY)}")
    cls = SCHEDULER_REGISTRY[key]
    n_train = int(spec.get("num_train_timesteps") or training_cfg.get("num_train_timesteps") or 1000)
    accepted = set(inspect.signature(cls.__init__).parameters) - {"self", "unused"}
    kwargs = {k: v for k, v in dict(spec.get("params", {})).items() if k in accepted}
    scheduler = cls(num_train_timesteps=n_train, **kwargs)
    n_infer = int(spec.get("num_inference_steps") or training_cfg.get("num_inference_steps") or n_train)
    return scheduler, n_infer


_OVERRIDES = {
    "ddpm": {"name": "ddpm"},
    "ddim": {"name": "ddim"},
    # the reference's aliases carry only solver_order + algorithm_type (`pipelines/utils.py:76-78`); with those alone
    # diffusers >= 0.26 raises "`final_sigmas_type` zero is not supported for `algorithm_type` dpmsolver. Please choose
    # `sigma_min` instead." - the aliases here add exactly that, so the names run
    "dpmsolver1": {"name": "dpm_multistep", "params": {"solver_order": 1, "algorithm_type": "dpmsolver",
                                                      "final_sigmas_type": "sigma_min"}},
    "dpmsolver2": {"name": "dpm_multistep", "params": {"solver_order": 2, "algorithm_type": "dpmsolver",
                                                      "final_sigmas_type": "sigma_min"}},
    "dpmsolver++": {"name": "dpm_multistep", "params": {"solver_order": 2, "algorithm_type": "dpmsolver++"}},
    "dpmsolversde": {"name": "dpm_sde"},
    "unipc": {"name": "unipc"},
    "flowmatch": {"name": "flow_match_euler"},
    "flow_match_euler": {"name": "flow_match_euler"},
}


def resolve_scheduler_override(name: Optional[str]) -> Optional[Dict]:
    """`--scheduler` alias -> scheduler config override (`pipelines/utils.py:65-90`)."""
    if not name:
        return None
    key = str(name).strip().lower()
    if not key:
        return None
    if key in _OVERRIDES:
        return {k: (dict(v) if isinstance(v, dict) else v) for k, v in _OVERRIDES[key].items()}
    if key in SCHEDULER_REGISTRY or key in _OUT_OF_SCOPE_SCHEDULERS:
        return {"name": key}
    raise ValueError(f"Unknown scheduler override '{name}'. Available: {', '.join(sorted(_OVERRIDES))}")


def _forward_model(model, inputs, timesteps, context_ca=None):
    outputs = model(inputs, timesteps, context_ca=context_ca) if context_ca is not None else model(inputs, timesteps)
    if isinstance(outputs, tuple):
        return outputs[0]
    return outputs.sample if hasattr(outputs, "sample") else outputs


def sync_if_cuda(device: torch.device) -> None:
    if device.type == "cuda" and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def _align_conditioning(condition, target_batch):
    if condition is None or condition.size(0) == target_batch:
        return condition
    reps = math.ceil(target_batch / condition.size(0))
    tiled = condition.repeat(reps, 1, 1, 1) if reps > 1 else condition
    return tiled[:target_batch]


def normalize_latent_conditioning(condition, mode):
    """Per-sample "standardize" / "minmax" normalisation of latent conditioning (`pipelines/utils.py:122-150`)."""
    if condition is None:
        return None
    kind = str(mode or "none").lower()
    if kind in {"none", "false", "off"}:
        return condition
    dims = tuple(range(2, condition.dim()))
    if kind == "standardize":
        return (condition - condition.mean(dim=dims, keepdim=True)) / (condition.std(dim=dims, keepdim=True) + 1e-6)
    if kind == "minmax":
        lo, hi = condition.amin(dim=dims, keepdim=True), condition.amax(dim=dims, keepdim=True)
        return (condition - lo) / (hi - lo + 1e-6)
    raise ValueError(f"Unknown latent_norm mode: {mode}")


# ------------------------------------------------------------------------------------------------------------------
def _model_signature(model) -> tuple:
    """(storage address, version) of every parameter and buffer: what a captured graph has baked in.  An optimiser
    step, an in-place `load_state_dict`, `.to()` or a re-seat of `p.data` changes it."""
    if not isinstance(model, torch.nn.Module):
        return ()
    sig = [(p.data_ptr(), p._version) for p in model.parameters()]
    sig.extend((b.data_ptr(), b._version) for b in model.buffers())
    return tuple(sig)


def _scheduler_key(scheduler) -> tuple:
    """Schedulers with the same class and config replay from the same graph (their per-run coefficient and timestep
    tables are reloaded into the static buffers before every run)."""
    cfg = getattr(scheduler, "config", None)
    items = tuple(sorted((k, repr(v)) for k, v in vars(cfg).items())) if cfg is not None else ()
    return (type(scheduler).__name__, items)


class GraphSampler:
    """[denoiser forward -> scheduler step -> cursor++] captured once as a CUDA graph and replayed per step.

    The sampler state x (fp32), the conditioning, the per-run coefficient table, the per-run timestep table and the
    int32 step cursor all live in static device buffers; nothing crosses the host between steps.  The captured kernels
    hold raw pointers to the packed weights and the fp32 parameters, so the graph is re-captured whenever the model's
    parameter signature (`_model_signature`) changes - sample -> train / load weights -> sample stays correct."""

    def __init__(self, model: BaseUNetND, scheduler: _SchedulerBase, shape, device, cond_shape=None, ctx_shape=None):
        self.model, self.scheduler = model, scheduler
        self.shape, self.device = tuple(shape), torch.device(device)
        with torch.cuda.device(self.device):
            self.x = torch.zeros(self.shape, dtype=torch.float32, device=device)
            self.cond = None if cond_shape is None else torch.zeros(cond_shape, dtype=torch.float32, device=device)
            # conditioning: "attention": the cross-attention context lives in a static buffer; its keys/values are
            # (re)computed in place by the attention modules before the replays (`precompute_context`), never per step
            self.ctx = None if ctx_shape is None else torch.zeros(ctx_shape, dtype=torch.float32, device=device)
            self.cursor = torch.zeros(1, dtype=torch.int32, device=device)
            self.state = scheduler.new_state(self.x)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.capacity = 0
        self.coef = None
        self.tvals = None
        self.launches_per_step = 0
        self.captures = 0
        self._sig: Optional[tuple] = None
        self._steps = 0

    def _one_step(self):
        kw = {} if self.ctx is None else {"context_ca": self.ctx}
        pred = self.model(self.x, None, context=self.cond, t_table=self.tvals, step_dev=self.cursor, **kw)
        self.scheduler.step_kernel(self.x, self.x, pred, self.coef, step_dev=self.cursor, state=self.state)
        ops.counter_add(self.cursor, 1)

    def _load_plan(self, scheduler, timesteps: torch.Tensor) -> int:
        coef, tvals = scheduler.run_plan(timesteps, self.device)
        n = tvals.numel()
        if self.coef is None or n > self.capacity:
            # >= 2 rows: the capture warm-up runs two steps (cursor 0 and 1) whatever the length of the plan
            self.capacity = max(n, self.capacity, 2)
            self.coef = torch.zeros((self.capacity, coef.shape[1]), dtype=torch.float32, device=self.device)
            self.tvals = torch.zeros((self.capacity,), dtype=torch.float32, device=self.device)
            self.graph = None
        self.coef[:n].copy_(coef)
        self.tvals[:n].copy_(tvals)
        return n

    # B*H*W up to which the step graph is captured with programmatic dependent launch: a step of such a problem is
    # ~150 kernels of 3-30 us whose launch latencies and prologues PDL overlaps (MNIST, batch 64: 925 vs 870
    # samples/s); on large problems the same attribute costs 1-2 %
    PDL_MAX_PIXELS = 1 << 17

    def _capture(self):
        pixels = self.shape[0] * self.shape[-2] * self.shape[-1]
        with ops.pdl(pixels <= self.PDL_MAX_PIXELS):
            self._capture_graph()

    def _capture_graph(self):
        # warm-up on a side stream (packs weights, sets kernel attributes, primes the allocator), state restored after
        saved_x = self.x.clone()
        saved_state = None if self.state is None else {k: v.clone() for k, v in self.state.items()}
        self.cursor.zero_()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            before = ops.launch_count()
            self._one_step()
            self.launches_per_step = ops.launch_count() - before
            self._one_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.cursor.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._one_step()
        self.graph = graph
        self.captures += 1
        self.x.copy_(saved_x)
        if saved_state is not None:
            for k, v in saved_state.items():
                self.state[k].copy_(v)

    def prepare(self, init: torch.Tensor, cond: Optional[torch.Tensor], timesteps: torch.Tensor,
                ctx: Optional[torch.Tensor] = None, scheduler=None) -> None:
        """Everything that is not the N-step loop: plan tables, input staging, context keys/values and - when the
        shapes, the table capacity or the model's parameters changed - the (re)capture of the step graph."""
        with torch.cuda.device(self.device):
            self._steps = self._load_plan(scheduler if scheduler is not None else self.scheduler, timesteps)
            self.x.copy_(init.to(device=self.device, dtype=torch.float32))
            if self.cond is not None:
                self.cond.copy_(cond.to(device=self.device, dtype=torch.float32))
            if self.ctx is not None:
                self.ctx.copy_(ctx.to(device=self.device, dtype=torch.float32))
                for m in self.model.modules():
                    if hasattr(m, "precompute_context"):
                        m.precompute_context(self.ctx)
            if self.state is not None:
                for v in self.state.values():
                    v.zero_()
            sig = _model_signature(self.model)
            if self.graph is None or sig != self._sig:
                self._capture()
                self._sig = sig
            self.cursor.zero_()

    def replay(self) -> torch.Tensor:
        with torch.cuda.device(self.device):
            for _ in range(self._steps):
                self.graph.replay()
            return self.x.clone()

    def run(self, init: torch.Tensor, cond: Optional[torch.Tensor], timesteps: torch.Tensor,
            ctx: Optional[torch.Tensor] = None, scheduler=None) -> torch.Tensor:
        self.prepare(init, cond, timesteps, ctx=ctx, scheduler=scheduler)
        return self.replay()


_GRAPH_CACHE: Dict[tuple, GraphSampler] = {}
_GRAPH_CACHE_MAX = 4


def _graph_sampler(model, scheduler, shape, device, cond_shape, ctx_shape=None) -> GraphSampler:
    key = (id(model), _scheduler_key(scheduler), tuple(shape), str(device),
           None if cond_shape is None else tuple(cond_shape), None if ctx_shape is None else tuple(ctx_shape))
    gs = _GRAPH_CACHE.get(key)
    if gs is not None and gs.model is not model:  # id() of a collected model reused by a new one
        gs = None
    if gs is None:
        while len(_GRAPH_CACHE) >= _GRAPH_CACHE_MAX:
            _GRAPH_CACHE.pop(next(iter(_GRAPH_CACHE)))
        gs = GraphSampler(model, scheduler, shape, device, cond_shape, ctx_shape)
        _GRAPH_CACHE[key] = gs
    return gs


@torch.no_grad()
def sample_with_scheduler(model: torch.nn.Module, scheduler, num_inference_steps: int,
                          sample_shape: Tuple[int, ...], device: torch.device,
                          conditioning_mode: Optional[str] = None,
                          conditioning_batch: Optional[torch.Tensor] = None, latent_norm: Optional[str] = None,
                          timing: Optional[dict] = None, start_step: Optional[int] = None,
                          last_n_steps: Optional[int] = None, init_sample: Optional[torch.Tensor] = None,
                          use_cuda_graph: bool = True) -> torch.Tensor:
    """Run the N-step sampling loop; same arguments and result as the reference's `sample_with_scheduler`."""
    device = torch.device(device)
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    if start_step is not None:
        start_step = int(start_step)
        if start_step < 0:
            raise ValueError("start_step must be >= 0.")
        timesteps = timesteps[timesteps <= start_step]
    if last_n_steps is not None:
        last_n_steps = int(last_n_steps)
        if last_n_steps <= 0:
            raise ValueError("last_n_steps must be > 0.")
        timesteps = timesteps[-last_n_steps:]
    if timesteps.numel() == 0:
        raise ValueError("No timesteps selected after applying start_step/last_n_steps.")

    current = init_sample.to(device) if init_sample is not None else torch.randn(sample_shape, device=device)
    cond = _align_conditioning(conditioning_batch, current.size(0))
    if cond is not None:
        cond = cond.to(device)
    if conditioning_mode == "attention":
        cond = normalize_latent_conditioning(cond, latent_norm)
    attention_ctx = cond if conditioning_mode == "attention" else None
    concat = conditioning_mode == "concatenate" and cond is not None

    graphable = (use_cuda_graph and device.type == "cuda" and isinstance(model, BaseUNetND)
                 and isinstance(scheduler, _SchedulerBase) and not model.training)
    if graphable:
        gs = _graph_sampler(model, scheduler, current.shape, device, cond.shape if concat else None,
                            None if attention_ctx is None else attention_ctx.shape)
        # staging and (re)capture stay outside the timed region: `model_seconds` is the N-step loop, as in the
        # reference (`pipelines/utils.py:211-217`)
        gs.prepare(current, cond if concat else None, timesteps, ctx=attention_ctx, scheduler=scheduler)
        sync_if_cuda(device)
        t0 = time.perf_counter()
        out = gs.replay()
        sync_if_cuda(device)
        if timing is not None:
            timing["model_seconds"] = timing.get("model_seconds", 0.0) + (time.perf_counter() - t0)
            timing["model_calls"] = timing.get("model_calls", 0) + int(timesteps.numel())
        return out

    # generic step-by-step loop (any callable model / any scheduler object), reference order of operations
    for t in timesteps:
        if concat:
            if isinstance(model, BaseUNetND):
                step_t = t if torch.is_tensor(t) else torch.as_tensor(t)
                step_t = step_t.to(current.device)
                if step_t.dim() == 0:
                    step_t = step_t.expand(current.size(0))
                sync_if_cuda(current.device)
                t0 = time.perf_counter()
                pred = model(current, step_t, context=cond)
            else:
                model_input = torch.cat([current, cond], dim=1)
                step_t = (t if torch.is_tensor(t) else torch.as_tensor(t)).to(current.device)
                if step_t.dim() == 0:
                    step_t = step_t.expand(current.size(0))
                sync_if_cuda(current.device)
                t0 = time.perf_counter()
                pred = _forward_model(model, model_input, step_t, context_ca=attention_ctx)
        else:
            step_t = (t if torch.is_tensor(t) else torch.as_tensor(t)).to(current.device)
            if step_t.dim() == 0:
                step_t = step_t.expand(current.size(0))
            sync_if_cuda(current.device)
            t0 = time.perf_counter()
            pred = _forward_model(model, current, step_t, context_ca=attention_ctx)
        sync_if_cuda(current.device)
        if timing is not None:
            timing["model_seconds"] = timing.get("model_seconds", 0.0) + (time.perf_counter() - t0)
            timing["model_calls"] = timing.get("model_calls", 0) + 1
        current = scheduler.step(pred, t, current).prev_sample
    return current
