"""B200-native schedulers: the duck-type `sample_with_scheduler` consumes (`src/pipelines/utils.py:163-220`,
`src/utils/model_utils/diffusion_utils.py:196-227`) — `set_timesteps`, `.timesteps`, `.step(pred, t, x).prev_sample`,
optional `.add_noise`, `.config.num_train_timesteps` — for the three samplers the north star names.

The reference binds these names to third-party `diffusers` classes (`pipelines/utils.py:13-30`); the arithmetic is
restated here from the published algorithms with diffusers' defaults and dtype discipline (float64 numpy schedules
cast to fp32; per-step coefficients evaluated as 0-dim fp32 tensors on the host, in diffusers' operation order).
Each `.step` is ONE fused elementwise kernel (K4) reading a device-resident coefficient table, so a whole sampling
run can be replayed from a CUDA graph with a device-side step cursor and no host round trip.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch

from .. import ops


class SchedulerOutput:
    __slots__ = ("prev_sample",)

    def __init__(self, prev_sample: torch.Tensor):
        self.prev_sample = prev_sample


def _lookup(schedule: torch.Tensor, timestep) -> int:
    """Position of `timestep` in `schedule` (second hit if the value repeats, like diffusers)."""
    value = timestep.to(schedule.device) if torch.is_tensor(timestep) else timestep
    hits = (schedule == value).nonzero()
    if len(hits) == 0:
        raise ValueError(f"timestep {timestep} is not in the scheduler's timesteps")
    return int(hits[1 if len(hits) > 1 else 0].item())


class _SchedulerBase:
    NCOEF = 1
    order = 1
    init_noise_sigma = 1.0

    def __init__(self):
        self._dev_tables: Dict[str, torch.Tensor] = {}
        self._coef_cpu: Optional[torch.Tensor] = None
        self._step_index: Optional[int] = None
        self.num_inference_steps: Optional[int] = None

    # -- device-resident coefficient table ---------------------------------------------------------------------
    def coef_table(self, device) -> torch.Tensor:
        key = str(device)
        tab = self._dev_tables.get(key)
        if tab is None:
            tab = self._coef_cpu.to(device=device, dtype=torch.float32).contiguous()
            self._dev_tables[key] = tab
        return tab

    def _reset_tables(self):
        self._dev_tables = {}
        self._step_index = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    # -- graph-replay support: rows / timestep values of a planned run, in run order -----------------------------
    def plan_rows(self, timesteps: torch.Tensor) -> list:
        raise NotImplementedError

    def run_plan(self, timesteps: torch.Tensor, device):
        rows = self.plan_rows(timesteps)
        coef = self.coef_table(device)[torch.as_tensor(rows, device=device)].contiguous()
        tvals = timesteps.to(device=device, dtype=torch.float32).contiguous()
        return coef, tvals

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        raise NotImplementedError

    def new_state(self, x: torch.Tensor):
        return None


# ------------------------------------------------------------------------------------------------------------------
class FlowMatchEulerDiscreteScheduler(_SchedulerBase):
    """x <- x + (sigma[i+1] - sigma[i]) * v.  shift = 1, no dynamic shifting (the reference's `params: {}`)."""

    NCOEF = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False,
                 **unused):
        super().__init__()
        if use_dynamic_shifting:
            raise NotImplementedError("fmdm_b200: dynamic shifting is outside the reference configs")
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), shift=float(shift),
                                      use_dynamic_shifting=False)
        T = self.config.num_train_timesteps
        grid = np.linspace(1, T, T, dtype=np.float32)[::-1].copy()
        sig = torch.from_numpy(grid).to(torch.float32) / T
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = sig * T
        self.sigmas = sig
        self.sigma_min = float(sig[-1])
        self.sigma_max = float(sig[0])
        self._build(sig)

    def _build(self, sig: torch.Tensor):
        full = torch.cat([sig, torch.zeros(1)]) if sig.numel() == self.timesteps.numel() else sig
        self._coef_cpu = (full[1:] - full[:-1]).reshape(-1, 1).contiguous()
        self._reset_tables()

    def set_timesteps(self, num_inference_steps: int, device=None):
        T, shift = self.config.num_train_timesteps, self.config.shift
        n = int(num_inference_steps)
        sig = np.linspace(self.sigma_max * T, self.sigma_min * T, n) / T  # float64
        sig = shift * sig / (1 + (shift - 1) * sig)
        sig = torch.from_numpy(sig).to(dtype=torch.float32)
        self.timesteps = sig * T
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.num_inference_steps = n
        self._build(self.sigmas)

    def plan_rows(self, timesteps):
        first = _lookup(self.timesteps, timesteps[0])
        return list(range(first, first + len(timesteps)))

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        ops.sched_flowmatch(x, pred, coef, step_host, x_out=x_out, step_dev=step_dev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, **unused) -> SchedulerOutput:
        if isinstance(timestep, int) or (torch.is_tensor(timestep) and not timestep.is_floating_point()):
            raise ValueError("Passing integer indices as timesteps to FlowMatchEulerDiscreteScheduler.step() is not "
                             "supported; pass one of `scheduler.timesteps`.")
        if self._step_index is None:
            self._step_index = _lookup(self.timesteps, timestep)
        x = sample.to(torch.float32).contiguous()
        v = model_output.to(torch.float32).contiguous()
        out = ops.sched_flowmatch(x, v, self.coef_table(x.device), self._step_index)
        self._step_index += 1
        return SchedulerOutput(out.to(model_output.dtype))


# ------------------------------------------------------------------------------------------------------------------
def _alphas_cumprod(T: int, beta_start: float, beta_end: float, beta_schedule: str) -> torch.Tensor:
    if beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    elif beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
    else:
        raise NotImplementedError(f"fmdm_b200: beta_schedule '{beta_schedule}' is not supported")
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMScheduler(_SchedulerBase):
    """Deterministic DDIM (eta = 0), epsilon prediction, clip_sample to +-clip_sample_range, "leading" spacing."""

    NCOEF = 4

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", clip_sample: bool = True, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", clip_sample_range: float = 1.0,
                 timestep_spacing: str = "leading", **unused):
        super().__init__()
        if prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("fmdm_b200 DDIM: only epsilon prediction with leading spacing")
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=bool(clip_sample),
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=int(steps_offset),
                                      prediction_type=prediction_type, clip_sample_range=float(clip_sample_range),
                                      timestep_spacing=timestep_spacing)
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(T, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        n = int(num_inference_steps)
        if n > T:
            raise ValueError(f"`num_inference_steps`: {n} cannot be larger than `num_train_timesteps`: {T}")
        self.num_inference_steps = n
        ratio = T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)
        rows = []
        for t in ts.tolist():
            prev_t = t - ratio
            a_t = self.alphas_cumprod[t]
            a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
            b_t = 1 - a_t
            b_prev = 1 - a_prev
            variance = (b_prev / b_t) * (1 - a_t / a_prev)
            std = 0.0 * variance ** 0.5
            rows.append(torch.stack([b_t ** 0.5, a_t ** 0.5, a_prev ** 0.5, (1 - a_prev - std ** 2) ** 0.5]))
        self._coef_cpu = torch.stack(rows).to(torch.float32).contiguous()
        self._reset_tables()

    def plan_rows(self, timesteps):
        return [_lookup(self.timesteps, t) for t in timesteps]

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        ops.sched_ddim(x, pred, coef, step_host, self.config.clip_sample, self.config.clip_sample_range, x_out=x_out,
                       step_dev=step_dev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             **unused) -> SchedulerOutput:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', run 'set_timesteps' first")
        if eta != 0.0:
            raise NotImplementedError("fmdm_b200 DDIM: eta must be 0")
        row = _lookup(self.timesteps, int(timestep))
        x = sample.to(torch.float32).contiguous()
        e = model_output.to(torch.float32).contiguous()
        out = ops.sched_ddim(x, e, self.coef_table(x.device), row, self.config.clip_sample,
                             self.config.clip_sample_range)
        return SchedulerOutput(out.to(model_output.dtype))

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(torch.float32)
        idx = timesteps.to("cpu", torch.int64).flatten()
        a = (ac[idx] ** 0.5).to(original_samples.device).contiguous()
        b = ((1 - ac[idx]) ** 0.5).to(original_samples.device).contiguous()
        return ops.sched_add_noise(original_samples, noise, a, b).to(original_samples.dtype)


# ------------------------------------------------------------------------------------------------------------------
class DDPMScheduler(_SchedulerBase):
    """Ancestral DDPM sampling (the reference's default scheduler name, `pipelines/utils.py:46`): epsilon prediction,
    variance_type "fixed_small", clip_sample to +-clip_sample_range, "leading" spacing.  The Gaussian noise of each
    step is drawn with torch's generator (as diffusers' `randn_tensor` does) into a static buffer, so the step is
    still one fused kernel and the run replays from a CUDA graph."""

    NCOEF = 8

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", variance_type: str = "fixed_small", clip_sample: bool = True,
                 prediction_type: str = "epsilon", clip_sample_range: float = 1.0, timestep_spacing: str = "leading",
                 steps_offset: int = 0, **unused):
        super().__init__()
        if prediction_type != "epsilon" or timestep_spacing != "leading" or variance_type != "fixed_small":
            raise NotImplementedError("fmdm_b200 DDPM: only epsilon prediction, leading spacing, fixed_small variance")
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, variance_type=variance_type,
                                      clip_sample=bool(clip_sample), prediction_type=prediction_type,
                                      clip_sample_range=float(clip_sample_range), timestep_spacing=timestep_spacing,
                                      steps_offset=int(steps_offset))
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(T, beta_start, beta_end, beta_schedule)
        self.one = torch.tensor(1.0)
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))
        self.set_timesteps(T)

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        n = int(num_inference_steps)
        if n > T:
            raise ValueError(f"`num_inference_steps`: {n} cannot be larger than `num_train_timesteps`: {T}")
        self.num_inference_steps = n
        ratio = T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)
        rows = []
        zero = torch.tensor(0.0)
        for t in ts.tolist():
            prev_t = t - ratio
            a_t = self.alphas_cumprod[t]
            a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
            b_t = 1 - a_t
            b_prev = 1 - a_prev
            cur_alpha = a_t / a_prev
            cur_beta = 1 - cur_alpha
            c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
            c_xt = cur_alpha ** 0.5 * b_prev / b_t
            variance = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
            sigma = variance ** 0.5 if t > 0 else zero
            rows.append(torch.stack([b_t ** 0.5, a_t ** 0.5, c_x0, c_xt, sigma, zero, zero, zero]))
        self._coef_cpu = torch.stack(rows).to(torch.float32).contiguous()
        self._reset_tables()

    def plan_rows(self, timesteps):
        return [_lookup(self.timesteps, t) for t in timesteps]

    def new_state(self, x: torch.Tensor):
        return {"noise": torch.zeros_like(x)}

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        state["noise"].normal_()  # graph-safe: torch registers the generator's philox offset with the capture
        ops.sched_ddpm(x, pred, state["noise"], coef, step_host, self.config.clip_sample,
                       self.config.clip_sample_range, x_out=x_out, step_dev=step_dev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             noise: Optional[torch.Tensor] = None, **unused) -> SchedulerOutput:
        """`noise` (optional, same shape): the step's standard-normal draw, for seeded parity runs; otherwise drawn
        from `generator` / the global generator on the sample's device."""
        row = _lookup(self.timesteps, int(timestep))
        x = sample.to(torch.float32).contiguous()
        e = model_output.to(torch.float32).contiguous()
        if noise is None:
            noise = torch.randn(x.shape, generator=generator, device=x.device, dtype=torch.float32)
        out = ops.sched_ddpm(x, e, noise.to(torch.float32).contiguous(), self.coef_table(x.device), row,
                             self.config.clip_sample, self.config.clip_sample_range)
        return SchedulerOutput(out.to(model_output.dtype))

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(torch.float32)
        idx = timesteps.to("cpu", torch.int64).flatten()
        a = (ac[idx] ** 0.5).to(original_samples.device).contiguous()
        b = ((1 - ac[idx]) ** 0.5).to(original_samples.device).contiguous()
        return ops.sched_add_noise(original_samples, noise, a, b).to(original_samples.dtype)


# ------------------------------------------------------------------------------------------------------------------
class DPMSolverMultistepScheduler(_SchedulerBase):
    """DPM-Solver multistep, order <= 2, midpoint, "linspace" spacing - what the `--scheduler dpmsolver++ / dpmsolver1 /
    dpmsolver2` aliases build (`pipelines/utils.py:76-79`): algorithm_type "dpmsolver++" (data prediction, final sigma 0)
    or "dpmsolver" (noise prediction).  diffusers rejects "dpmsolver" with final_sigmas_type="zero" at construction
    (sigma_t = 0 times exp(h) = inf); the same ValueError is raised here, and `final_sigmas_type="sigma_min"` - which the
    dpmsolver1/2 aliases of this package add, as that error message instructs - is the runnable form."""

    NCOEF = 8

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", solver_order: int = 2, prediction_type: str = "epsilon",
                 algorithm_type: str = "dpmsolver++", solver_type: str = "midpoint", lower_order_final: bool = True,
                 euler_at_final: bool = False, final_sigmas_type: str = "zero", timestep_spacing: str = "linspace",
                 **unused):
        super().__init__()
        if algorithm_type not in ("dpmsolver++", "dpmsolver") or solver_type != "midpoint" \
                or prediction_type != "epsilon" or solver_order not in (1, 2) or timestep_spacing != "linspace" \
                or euler_at_final:
            raise NotImplementedError("fmdm_b200 DPM-Solver: dpmsolver++ / dpmsolver, midpoint, epsilon prediction, "
                                      "order <= 2, linspace spacing (the --scheduler dpmsolver++/dpmsolver1/dpmsolver2 "
                                      "aliases)")
        if algorithm_type != "dpmsolver++" and final_sigmas_type == "zero":
            raise ValueError(f"`final_sigmas_type` {final_sigmas_type} is not supported for `algorithm_type` "
                             f"{algorithm_type}. Please choose `sigma_min` instead.")
        if final_sigmas_type not in ("zero", "sigma_min"):
            raise ValueError(f"`final_sigmas_type` must be one of 'zero', or 'sigma_min', but got {final_sigmas_type}")
        self.order = int(solver_order)
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, solver_order=int(solver_order),
                                      prediction_type=prediction_type, algorithm_type=algorithm_type,
                                      solver_type=solver_type, lower_order_final=bool(lower_order_final),
                                      euler_at_final=False, final_sigmas_type=final_sigmas_type,
                                      timestep_spacing=timestep_spacing)
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(T, beta_start, beta_end, beta_schedule)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.timesteps = torch.from_numpy(np.linspace(0, T - 1, T, dtype=np.float32)[::-1].copy())
        self.lower_order_nums = 0
        self._state: Dict[str, torch.Tensor] = {}

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        n = int(num_inference_steps)
        ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
        if self.config.final_sigmas_type == "sigma_min":
            sigma_last = float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5)
        else:
            sigma_last = 0.0
        self.sigmas = torch.from_numpy(np.concatenate([sig, [sigma_last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(torch.int64)
        self.num_inference_steps = len(ts)
        self.lower_order_nums = 0
        self._state = {}
        raw = self.config.algorithm_type == "dpmsolver"
        # rows [0, n): first-order update at index i; rows [n, 2n): second-order update at index i
        L = len(ts)
        tab = torch.zeros((2 * L, self.NCOEF), dtype=torch.float32)
        lam = []
        for i in range(L + 1):
            a, s = self._alpha_sigma(self.sigmas[i])
            lam.append(torch.log(a) - torch.log(s))
        for i in range(L):
            a_s, s_s = self._alpha_sigma(self.sigmas[i])
            a_t, s_t = self._alpha_sigma(self.sigmas[i + 1])
            h = lam[i + 1] - lam[i]
            if raw:
                c1 = a_t / a_s
                c2 = s_t * (torch.exp(h) - 1.0)
            else:
                c1 = s_t / s_s
                c2 = a_t * (torch.exp(-h) - 1.0)
            for second in (0, 1):
                row = tab[second * L + i]
                row[0], row[1], row[2], row[3] = s_s, a_s, c1, c2
                row[7] = 1.0 if raw else 0.0
                if second and i >= 1:
                    h0 = lam[i] - lam[i - 1]
                    r0 = h0 / h
                    row[4] = 0.5 * c2
                    row[5] = 1.0 / r0
                    row[6] = 1.0
        self._coef_cpu = tab.contiguous()
        self._reset_tables()

    def _row(self, index: int, lower_order_nums: int) -> int:
        L = len(self.timesteps)
        final = (index == L - 1) and ((self.config.lower_order_final and L < 15)
                                      or self.config.final_sigmas_type == "zero")
        first = self.config.solver_order == 1 or lower_order_nums < 1 or final
        return index if first else L + index

    def plan_rows(self, timesteps):
        first = _lookup(self.timesteps, timesteps[0])
        return [self._row(first + k, min(k, self.config.solver_order)) for k in range(len(timesteps))]

    def new_state(self, x: torch.Tensor):
        return {"m": torch.zeros_like(x, dtype=torch.float32)}

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        m = state["m"]
        ops.sched_dpmpp2m(x, pred, m, coef, step_host, x_out=x_out, m_cur=m, step_dev=step_dev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, **unused) -> SchedulerOutput:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', run 'set_timesteps' first")
        if self._step_index is None:
            self._step_index = _lookup(self.timesteps, timestep)
        x = sample.to(torch.float32).contiguous()
        e = model_output.to(torch.float32).contiguous()
        if "m" not in self._state or self._state["m"].shape != x.shape or self._state["m"].device != x.device:
            self._state = self.new_state(x)
        row = self._row(self._step_index, self.lower_order_nums)
        m = self._state["m"]
        out, _ = ops.sched_dpmpp2m(x, e, m, self.coef_table(x.device), row, m_cur=m)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return SchedulerOutput(out.to(model_output.dtype))

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        idx = [_lookup(self.timesteps, t) for t in timesteps.to("cpu").flatten()]
        sigma = self.sigmas[idx].flatten()
        a, s = self._alpha_sigma(sigma)
        dev = original_samples.device
        return ops.sched_add_noise(original_samples, noise, a.to(dev).contiguous(), s.to(dev).contiguous()).to(
            original_samples.dtype)


# ------------------------------------------------------------------------------------------------------------------
class UniPCMultistepScheduler(_SchedulerBase):
    """UniPC (`--scheduler unipc`, `pipelines/utils.py:28,82`): bh2, data prediction, order <= 2, lower_order_final,
    "linspace" spacing, final sigma 0 - diffusers' defaults.  One fused kernel per step (`fm_sched_unipc_f32`): the
    corrector of the incoming sample, the predictor and the history shift; per-step scalars (incl. the 2x2 solve of the
    order-2 corrector weights) are evaluated on the host as 0-dim fp32 tensors when the timesteps are set."""

    NCOEF = 16
    _VARIANTS = ((1, 0), (2, 0), (1, 1), (2, 1), (1, 2), (2, 2))  # (predictor order, corrector order; 0 = no corrector)

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", solver_order: int = 2, prediction_type: str = "epsilon",
                 predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector=(), timestep_spacing: str = "linspace", final_sigmas_type: str = "zero", **unused):
        super().__init__()
        if prediction_type != "epsilon" or not predict_x0 or solver_type != "bh2" or solver_order not in (1, 2) \
                or list(disable_corrector) or timestep_spacing != "linspace" or final_sigmas_type != "zero":
            raise NotImplementedError("fmdm_b200 UniPC: bh2 / predict_x0 / epsilon / order <= 2 / linspace spacing / "
                                      "final sigma zero (the --scheduler unipc alias)")
        self.order = int(solver_order)
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, solver_order=int(solver_order),
                                      prediction_type=prediction_type, predict_x0=True, solver_type=solver_type,
                                      lower_order_final=bool(lower_order_final), disable_corrector=[],
                                      timestep_spacing=timestep_spacing, final_sigmas_type=final_sigmas_type)
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(T, beta_start, beta_end, beta_schedule)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.timesteps = torch.from_numpy(np.linspace(0, T - 1, T, dtype=np.float32)[::-1].copy())
        self._reset_run()

    _alpha_sigma = staticmethod(DPMSolverMultistepScheduler._alpha_sigma)

    def _reset_run(self):
        self.lower_order_nums = 0
        self.this_order = None
        self._have_last = False
        self._state: Dict[str, torch.Tensor] = {}

    def _lam(self, i: int):
        a, s = self._alpha_sigma(self.sigmas[i])
        return torch.log(a) - torch.log(s)

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        n = int(num_inference_steps)
        ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(torch.int64)
        self.num_inference_steps = len(ts)
        self._reset_run()
        L = len(ts)
        tab = torch.zeros((len(self._VARIANTS) * L, self.NCOEF), dtype=torch.float32)
        for i in range(L):
            a_c, s_c = self._alpha_sigma(self.sigmas[i])           # conversion / corrector target (sigmas[i])
            a_n, s_n = self._alpha_sigma(self.sigmas[i + 1])       # predictor target
            lam_c = torch.log(a_c) - torch.log(s_c)
            h_p = (torch.log(a_n) - torch.log(s_n)) - lam_c
            hh_p = -h_p
            phi1_p, B_p = torch.expm1(hh_p), torch.expm1(hh_p)
            pred = {"pa": s_n / s_c, "pb": a_n * phi1_p, "pc": a_n * B_p}
            if i >= 1:
                pred["rk"] = (self._lam(i - 1) - lam_c) / h_p
                a_0, s_0 = self._alpha_sigma(self.sigmas[i - 1])
                lam_0 = torch.log(a_0) - torch.log(s_0)
                h_c = lam_c - lam_0
                hh_c = -h_c
                phi1_c, B_c = torch.expm1(hh_c), torch.expm1(hh_c)
                corr = {"ca": s_c / s_0, "cb": a_c * phi1_c, "cc": a_c * B_c}
                if i >= 2:
                    rk = (self._lam(i - 2) - lam_0) / h_c
                    rks = torch.stack([rk, torch.tensor(1.0)])
                    R, b = [], []
                    h_phi_k = phi1_c / hh_c - 1
                    factorial_i = 1
                    for k in range(1, 3):
                        R.append(torch.pow(rks, k - 1))
                        b.append(h_phi_k * factorial_i / B_c)
                        factorial_i *= k + 1
                        h_phi_k = h_phi_k / hh_c - 1 / factorial_i
                    corr["rho2"] = torch.linalg.solve(torch.stack(R), torch.stack(b))
                    corr["rk"] = rk
            for v, (po, co) in enumerate(self._VARIANTS):
                if (po == 2 and i < 1) or (co >= 1 and i < 1) or (co == 2 and i < 2):
                    continue
                row = tab[v * L + i]
                row[0], row[1] = s_c, a_c
                if co:
                    row[2], row[3], row[4], row[5] = 1.0, corr["ca"], corr["cb"], corr["cc"]
                    if co == 2:
                        row[6], row[7], row[8], row[9] = 1.0, corr["rk"], corr["rho2"][0], corr["rho2"][1]
                    else:
                        row[9] = 0.5
                row[10], row[11], row[12] = pred["pa"], pred["pb"], pred["pc"]
                if po == 2:
                    row[13], row[14], row[15] = 1.0, pred["rk"], 0.5
        self._coef_cpu = tab.contiguous()
        self._reset_tables()

    def _pred_order(self, index: int, lower_order_nums: int) -> int:
        L = len(self.timesteps)
        this_order = min(self.config.solver_order, L - index) if self.config.lower_order_final \
            else self.config.solver_order
        return min(this_order, lower_order_nums + 1)

    def _row(self, index: int, pred_order: int, corr_order: int) -> int:
        return self._VARIANTS.index((pred_order, corr_order)) * len(self.timesteps) + index

    def plan_rows(self, timesteps):
        first = _lookup(self.timesteps, timesteps[0])
        rows, prev_order = [], 0
        for k in range(len(timesteps)):
            i = first + k
            po = self._pred_order(i, min(k, self.config.solver_order))
            co = prev_order if (i > 0 and k > 0) else 0   # a run that starts mid-schedule has no last sample yet
            rows.append(self._row(i, po, co))
            prev_order = po
        return rows

    def new_state(self, x: torch.Tensor):
        z = torch.zeros_like(x, dtype=torch.float32)
        return {"last": z, "m1": z.clone(), "m2": z.clone()}

    def step_kernel(self, x_out, x, pred, coef, step_host=0, step_dev=None, state=None):
        ops.sched_unipc(x, pred, state["last"], state["m1"], state["m2"], coef, step_host, x_out=x_out,
                        step_dev=step_dev)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, **unused) -> SchedulerOutput:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', run 'set_timesteps' first")
        if self._step_index is None:
            self._step_index = _lookup(self.timesteps, timestep)
        x = sample.to(torch.float32).contiguous()
        e = model_output.to(torch.float32).contiguous()
        if "m1" not in self._state or self._state["m1"].shape != x.shape or self._state["m1"].device != x.device:
            self._state = self.new_state(x)
        i = self._step_index
        co = self.this_order if (i > 0 and self._have_last) else 0
        po = self._pred_order(i, self.lower_order_nums)
        st = self._state
        out = ops.sched_unipc(x, e, st["last"], st["m1"], st["m2"], self.coef_table(x.device), self._row(i, po, co))
        self.this_order = po
        self._have_last = True
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return SchedulerOutput(out.to(model_output.dtype))

    add_noise = DPMSolverMultistepScheduler.add_noise
