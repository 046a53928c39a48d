"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the third-party schedulers the reference binds at
`/root/reference/src/pipelines/utils.py:13-30` and steps at `:218`:

    diffusers.FlowMatchEulerDiscreteScheduler, diffusers.DDIMScheduler, diffusers.DPMSolverMultistepScheduler
    (dpmsolver++ and dpmsolver algorithm types), diffusers.UniPCMultistepScheduler, diffusers.DDPMScheduler,
    diffusers.DPMSolverSDEScheduler

The arithmetic lives in `diffusers` (requirement `diffusers>=0.24.0`, `/root/reference/requirements.txt:18`, not
pinned, not vendored, NOT installed in this image or on the GPU box, no source copy on disk).  What follows restates
the published algorithms with the defaults `build_scheduler` instantiates (`pipelines/utils.py:53-60`: only
`num_train_timesteps` + the config's `params`), in plain torch-CPU / numpy, reproducing diffusers' dtype discipline
(float64 numpy linspace -> fp32 tensors; per-step coefficients as 0-dim fp32 tensors; `sample` upcast to fp32).

PARITY UNPINNED at this boundary: the reference's tests hold no golden vector for any scheduler and
diffusers cannot be run here.  The restatement is pinned only by analytic known-answer tests
(tests/test_oracle_schedulers.py: point-mass data => exact recovery; sum(dt) == -1; closed-form DDIM inversion).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class StepOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


def _index_for_timestep(schedule: torch.Tensor, timestep) -> int:
    """diffusers `index_for_timestep`: first match, or the second one when the value repeats."""
    t = timestep.to(schedule.device) if torch.is_tensor(timestep) else timestep
    hits = (schedule == t).nonzero()
    pos = 1 if len(hits) > 1 else 0
    return int(hits[pos].item())


# ------------------------------------------------------------------------------------------------------------------
class FlowMatchEulerOracle:
    """FlowMatchEulerDiscreteScheduler, shift = 1, no dynamic shifting (reference config `params: {}`)."""

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0):
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), shift=float(shift))
        T = self.config.num_train_timesteps
        ts = np.linspace(1, T, T, dtype=np.float32)[::-1].copy()
        sig = torch.from_numpy(ts).to(torch.float32) / T
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = sig * T
        self.sigmas = sig
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        shift = self.config.shift
        ts = np.linspace(self.sigma_max * T, self.sigma_min * T, int(num_inference_steps))  # float64
        sig = ts / T
        sig = shift * sig / (1 + (shift - 1) * sig)
        sig = torch.from_numpy(sig).to(dtype=torch.float32)
        self.timesteps = sig * T
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.num_inference_steps = int(num_inference_steps)
        self._step_index = None

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> StepOutput:
        if isinstance(timestep, int) or (torch.is_tensor(timestep) and not timestep.is_floating_point()):
            raise ValueError("FlowMatchEuler.step needs one of scheduler.timesteps (float), not an integer index")
        if self._step_index is None:
            self._step_index = _index_for_timestep(self.timesteps, timestep)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        sigma_next = self.sigmas[self._step_index + 1]
        prev = sample + (sigma_next - sigma) * model_output
        self._step_index += 1
        return StepOutput(prev.to(model_output.dtype))


# ------------------------------------------------------------------------------------------------------------------
def _linear_alphas_cumprod(T: int, beta_start: float, beta_end: float) -> torch.Tensor:
    betas = torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMOracle:
    """DDIMScheduler defaults: linear betas, clip_sample=True (range 1), set_alpha_to_one=True, steps_offset=0,
    timestep_spacing="leading", prediction_type="epsilon"; step(eta=0, use_clipped_model_output=False)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 clip_sample: bool = True, clip_sample_range: float = 1.0, set_alpha_to_one: bool = True,
                 steps_offset: int = 0):
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type="epsilon", timestep_spacing="leading")
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _linear_alphas_cumprod(T, beta_start, beta_end)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = int(num_inference_steps)
        ratio = T // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> StepOutput:
        T = self.config.num_train_timesteps
        t = int(timestep)
        prev_t = t - T // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        eps = model_output
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = 0.0 * variance ** 0.5  # eta = 0
        direction = (1 - a_prev - std ** 2) ** 0.5 * eps
        prev = a_prev ** 0.5 * x0 + direction
        return StepOutput(prev, x0)

    def add_noise(self, original: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(dtype=original.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.dim() < original.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original + sb * noise


# ------------------------------------------------------------------------------------------------------------------
class DDPMOracle:
    """diffusers.DDPMScheduler (the reference's default scheduler name, `pipelines/utils.py:13-30,46`) with its
    defaults: linear betas, variance_type="fixed_small", clip_sample=True (range 1), prediction_type="epsilon",
    timestep_spacing="leading", thresholding off.  Restated from the published algorithm (Ho et al. 2020, eq. 7 for
    the posterior mean, beta-tilde for the variance) in diffusers' operation order; PARITY UNPINNED like the others."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 clip_sample: bool = True, clip_sample_range: float = 1.0, steps_offset: int = 0):
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                      steps_offset=steps_offset, prediction_type="epsilon",
                                      variance_type="fixed_small", timestep_spacing="leading")
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _linear_alphas_cumprod(T, beta_start, beta_end)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = T
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = int(num_inference_steps)
        ratio = T // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def _variance(self, t: int, prev_t: int):
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        cur_beta = 1 - a_t / a_prev
        return torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, noise=None, generator=None) -> StepOutput:
        T = self.config.num_train_timesteps
        t = int(timestep)
        prev_t = t - T // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_prev / b_t
        prev = c_x0 * x0 + c_xt * sample
        variance = 0
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            variance = (self._variance(t, prev_t) ** 0.5) * noise
        prev = prev + variance
        return StepOutput(prev, x0)

    add_noise = DDIMOracle.add_noise


# ------------------------------------------------------------------------------------------------------------------
class DPMSolverPPOracle:
    """DPMSolverMultistepScheduler as the reference's aliases build it (`pipelines/utils.py:76-79`):
      `dpmsolver++`  solver_order=2, algorithm_type="dpmsolver++"  (data prediction)
      `dpmsolver1/2` solver_order=1/2, algorithm_type="dpmsolver"  (noise prediction)
    with diffusers' other defaults: solver_type="midpoint", lower_order_final=True, euler_at_final=False,
    timestep_spacing="linspace", epsilon prediction, no thresholding / Karras sigmas, final_sigmas_type="zero".

    diffusers (>= 0.26) REJECTS algorithm_type="dpmsolver" with final_sigmas_type="zero" at construction ("`final_sigmas_
    type` zero is not supported for `algorithm_type` dpmsolver. Please choose `sigma_min` instead."): the update would
    multiply sigma_t = 0 by exp(h) = inf.  The restatement raises the same ValueError; `final_sigmas_type="sigma_min"`
    (last sigma = sigma of training step 0) is the runnable form of the noise-prediction solver."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 solver_order: int = 2, algorithm_type: str = "dpmsolver++", final_sigmas_type: str = "zero",
                 lower_order_final: bool = True):
        if algorithm_type not in ("dpmsolver++", "dpmsolver") or solver_order not in (1, 2):
            raise NotImplementedError("oracle restates dpmsolver++ / dpmsolver of order 1 or 2 only")
        if algorithm_type != "dpmsolver++" and final_sigmas_type == "zero":
            raise ValueError(f"`final_sigmas_type` {final_sigmas_type} is not supported for `algorithm_type` "
                             f"{algorithm_type}. Please choose `sigma_min` instead.")
        if final_sigmas_type not in ("zero", "sigma_min"):
            raise ValueError(f"`final_sigmas_type` must be one of 'zero', or 'sigma_min', but got {final_sigmas_type}")
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, solver_order=solver_order, algorithm_type=algorithm_type,
                                      solver_type="midpoint", lower_order_final=bool(lower_order_final),
                                      euler_at_final=False, final_sigmas_type=final_sigmas_type,
                                      prediction_type="epsilon", timestep_spacing="linspace")
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _linear_alphas_cumprod(T, beta_start, beta_end)
        self.init_noise_sigma = 1.0
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.timesteps = torch.from_numpy(np.linspace(0, T - 1, T, dtype=np.float32)[::-1].copy())
        self.num_inference_steps = None
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        N = int(num_inference_steps)
        last_timestep = T  # lambda_min_clipped = -inf => nothing clipped
        ts = np.linspace(0, last_timestep - 1, N + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
        if self.config.final_sigmas_type == "sigma_min":
            sigma_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        else:
            sigma_last = 0
        sig = np.concatenate([sig, [sigma_last]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig)
        self.timesteps = torch.from_numpy(ts).to(dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        sigma_t = sigma * alpha_t
        return alpha_t, sigma_t

    def _first_order(self, m0, sample):
        i = self._step_index
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        alpha_s, sigma_s = self._alpha_sigma(self.sigmas[i])
        lam_t = torch.log(alpha_t) - torch.log(sigma_t)
        lam_s = torch.log(alpha_s) - torch.log(sigma_s)
        h = lam_t - lam_s
        if self.config.algorithm_type == "dpmsolver++":
            return (sigma_t / sigma_s) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * m0
        return (alpha_t / alpha_s) * sample - (sigma_t * (torch.exp(h) - 1.0)) * m0

    def _second_order(self, sample):
        i = self._step_index
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        alpha_s1, sigma_s1 = self._alpha_sigma(self.sigmas[i - 1])
        lam_t = torch.log(alpha_t) - torch.log(sigma_t)
        lam_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        lam_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
        m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
        h, h_0 = lam_t - lam_s0, lam_s0 - lam_s1
        r0 = h_0 / h
        D0, D1 = m0, (1.0 / r0) * (m0 - m1)
        if self.config.algorithm_type == "dpmsolver++":
            return ((sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                    - 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * D1)
        return ((alpha_t / alpha_s0) * sample - (sigma_t * (torch.exp(h) - 1.0)) * D0
                - 0.5 * (sigma_t * (torch.exp(h) - 1.0)) * D1)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> StepOutput:
        if self._step_index is None:
            self._step_index = _index_for_timestep(self.timesteps, timestep)
        n = len(self.timesteps)
        lower_order_final = (self._step_index == n - 1) and (
            self.config.euler_at_final or (self.config.lower_order_final and n < 15)
            or self.config.final_sigmas_type == "zero")
        lower_order_second = (self._step_index == n - 2) and self.config.lower_order_final and n < 15
        if self.config.algorithm_type == "dpmsolver++":  # data prediction from epsilon
            sigma = self.sigmas[self._step_index]
            alpha_t, sigma_t = self._alpha_sigma(sigma)
            converted = (sample - sigma_t * model_output) / alpha_t
        else:                                            # "dpmsolver": the solver integrates epsilon itself
            converted = model_output
        for k in range(self.config.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = converted
        sample = sample.to(torch.float32)
        if self.config.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = self._first_order(converted, sample)
        else:  # solver_order == 2 (lower_order_second also lands here for order 2)
            prev = self._second_order(sample)
        del lower_order_second
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return StepOutput(prev.to(model_output.dtype))

    def add_noise(self, original: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        idx = [_index_for_timestep(self.timesteps, t) for t in timesteps]
        sigma = self.sigmas[idx].flatten()
        while sigma.dim() < original.dim():
            sigma = sigma.unsqueeze(-1)
        alpha_t, sigma_t = self._alpha_sigma(sigma)
        return alpha_t * original + sigma_t * noise


# ------------------------------------------------------------------------------------------------------------------
class UniPCOracle:
    """diffusers.UniPCMultistepScheduler (`--scheduler unipc`, `pipelines/utils.py:28,82`) with its defaults:
    solver_order=2, predict_x0=True, solver_type="bh2", lower_order_final=True, disable_corrector=[], solver_p=None,
    timestep_spacing="linspace", final_sigmas_type="zero", epsilon prediction, no thresholding / Karras sigmas.
    Restated from the published UniPC algorithm (Zhao et al. 2023: B(h) = expm1(-h) predictor UniP-p and corrector UniC-p
    on the data prediction) in diffusers' operation order; every step first corrects the incoming sample with the new
    model output (order = the previous step's), then predicts.  PARITY UNPINNED like the others."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 solver_order: int = 2, lower_order_final: bool = True):
        if solver_order not in (1, 2):
            raise NotImplementedError("oracle restates UniPC of order 1 or 2 only")
        self.config = SimpleNamespace(num_train_timesteps=int(num_train_timesteps), beta_start=beta_start,
                                      beta_end=beta_end, solver_order=solver_order, predict_x0=True, solver_type="bh2",
                                      lower_order_final=bool(lower_order_final), disable_corrector=[],
                                      prediction_type="epsilon", timestep_spacing="linspace", final_sigmas_type="zero")
        T = self.config.num_train_timesteps
        self.alphas_cumprod = _linear_alphas_cumprod(T, beta_start, beta_end)
        self.init_noise_sigma = 1.0
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.timesteps = torch.from_numpy(np.linspace(0, T - 1, T, dtype=np.float32)[::-1].copy())
        self.num_inference_steps = None
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        N = int(num_inference_steps)
        ts = np.linspace(0, T - 1, N + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
        sig = np.concatenate([sig, [0]]).astype(np.float32)  # final_sigmas_type == "zero"
        self.sigmas = torch.from_numpy(sig)
        self.timesteps = torch.from_numpy(ts).to(dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None

    _alpha_sigma = staticmethod(DPMSolverPPOracle._alpha_sigma)

    def _lambda(self, sigma):
        alpha_t, sigma_t = self._alpha_sigma(sigma)
        return torch.log(alpha_t) - torch.log(sigma_t)

    @staticmethod
    def _bh2(h):
        """(h_phi_1, B_h, R-matrix builder inputs) for predict_x0 / bh2: hh = -h."""
        hh = -h
        h_phi_1 = torch.expm1(hh)
        return hh, h_phi_1, torch.expm1(hh)

    def _rhos(self, rks, hh, h_phi_1, B_h, order):
        """R (order x order) and b (order) of the UniPC linear system."""
        R, b = [], []
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        return torch.stack(R), torch.stack(b)

    def _predict(self, sample, order):
        i = self._step_index
        m0 = self.model_outputs[-1]
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        lam_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = (torch.log(alpha_t) - torch.log(sigma_t)) - lam_s0
        D1 = None
        if order == 2:
            rk = (self._lambda(self.sigmas[i - 1]) - lam_s0) / h
            D1 = (self.model_outputs[-2] - m0) / rk
        hh, h_phi_1, B_h = self._bh2(h)
        x_t_ = sigma_t / sigma_s0 * sample - alpha_t * h_phi_1 * m0
        pred_res = torch.tensor(0.5) * D1 if D1 is not None else 0  # order 2: rhos_p = [0.5] (diffusers' shortcut)
        return x_t_ - alpha_t * B_h * pred_res

    def _correct(self, model_t, last_sample, order):
        i = self._step_index
        m0 = self.model_outputs[-1]  # the previous step's data prediction (history not shifted yet)
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i - 1])
        lam_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = (torch.log(alpha_t) - torch.log(sigma_t)) - lam_s0
        rks, D1 = [], None
        if order == 2:
            rk = (self._lambda(self.sigmas[i - 2]) - lam_s0) / h
            rks.append(rk)
            D1 = (self.model_outputs[-2] - m0) / rk
        rks.append(torch.tensor(1.0))
        rks = torch.stack(rks)
        hh, h_phi_1, B_h = self._bh2(h)
        if order == 1:
            rhos_c = torch.tensor([0.5])
        else:
            R, b = self._rhos(rks, hh, h_phi_1, B_h, order)
            rhos_c = torch.linalg.solve(R, b)
        x_t_ = sigma_t / sigma_s0 * last_sample - alpha_t * h_phi_1 * m0
        corr_res = rhos_c[0] * D1 if D1 is not None else 0
        D1_t = model_t - m0
        return x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> StepOutput:
        if self._step_index is None:
            self._step_index = _index_for_timestep(self.timesteps, timestep)
        i = self._step_index
        use_corrector = i > 0 and self.last_sample is not None
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
        converted = (sample - sigma_t * model_output) / alpha_t
        if use_corrector:
            sample = self._correct(converted, self.last_sample, self.this_order)
        for k in range(self.config.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = converted
        if self.config.lower_order_final:
            this_order = min(self.config.solver_order, len(self.timesteps) - i)
        else:
            this_order = self.config.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return StepOutput(prev.to(model_output.dtype))

    add_noise = DPMSolverPPOracle.add_noise


ORACLE_REGISTRY = {
    "flow_match_euler": FlowMatchEulerOracle,
    "flowmatch": FlowMatchEulerOracle,
    "ddim": DDIMOracle,
    "ddpm": DDPMOracle,
    "dpm_multistep": DPMSolverPPOracle,
    "unipc": UniPCOracle,
}
