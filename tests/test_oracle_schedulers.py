"""Pins the oracle's scheduler restatement with analytic known-answer tests (the reference holds no golden vector
for any scheduler and diffusers is not installable here: SURVEY.md §8c, "parity unpinned")."""
import math

import numpy as np
import pytest
import torch

from oracle.sampling import make_scheduler, sample_loop, select_timesteps
from oracle.schedulers import DDIMOracle, DPMSolverPPOracle, FlowMatchEulerOracle


def test_flowmatch_schedule_values():
    s = FlowMatchEulerOracle(1000)
    assert s.sigma_min == pytest.approx(0.0010000000474974513, abs=0) and s.sigma_max == 1.0
    s.set_timesteps(50)
    ts = s.timesteps
    assert ts.dtype == torch.float32 and ts.shape == (50,)
    assert float(ts[0]) == 1000.0 and float(ts[-1]) == 1.0
    assert float(ts[1]) == pytest.approx(979.6122, abs=1e-3) and float(ts[-2]) == pytest.approx(21.3877, abs=1e-3)
    assert s.sigmas.shape == (51,) and float(s.sigmas[-1]) == 0.0
    dts = s.sigmas[1:] - s.sigmas[:-1]
    assert float(dts.double().sum()) == pytest.approx(-1.0, abs=1e-6)


def test_flowmatch_point_mass_recovery():
    # data distribution = point mass at x0: the exact velocity field is v(x, t) = (x - x0) / sigma_t ... with the
    # reference's training target (noise - x0) the ODE x_t = (1-s) x0 + s*noise has dx/ds = noise - x0 = (x - x0)/s
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 1.0]]]])
    s = FlowMatchEulerOracle(1000)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(x0.shape, generator=g)
    state = {}

    def model(x, t):
        sigma = (t / 1000.0).view(-1, 1, 1, 1)
        return (x - x0) / sigma

    out = sample_loop(model, s, 50, noise)
    assert float((out - x0).abs().max()) < 1e-6


def test_flowmatch_rejects_integer_timesteps():
    s = FlowMatchEulerOracle(1000)
    s.set_timesteps(10)
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 5, torch.zeros(1))


def test_ddim_schedule_and_point_mass():
    s = DDIMOracle(1000, 1e-4, 0.02)
    s.set_timesteps(50)
    assert s.timesteps.dtype == torch.int64
    assert s.timesteps[:3].tolist() == [980, 960, 940] and int(s.timesteps[-1]) == 0
    # alphas_cumprod closed form check in float64
    betas = np.linspace(1e-4, 0.02, 1000)
    ac = np.cumprod(1 - betas)
    assert np.allclose(s.alphas_cumprod.numpy(), ac, rtol=2e-5)
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 0.9]]]])
    g = torch.Generator().manual_seed(1)
    a = s.alphas_cumprod[980]
    xT = a.sqrt() * x0 + (1 - a).sqrt() * torch.randn(x0.shape, generator=g)

    def model(x, t):  # exact epsilon for a point mass
        at = s.alphas_cumprod[t.long()].view(-1, 1, 1, 1)
        return (x - at.sqrt() * x0) / (1 - at).sqrt()

    out = sample_loop(model, s, 50, xT)
    assert float((out - x0).abs().max()) < 2e-5
    # clipping: a point mass outside [-1, 1] is clamped
    x0b = torch.full((1, 1, 2, 2), 1.7)
    xT = a.sqrt() * x0b + (1 - a).sqrt() * torch.randn(x0b.shape, generator=g)

    def model_b(x, t):
        at = s.alphas_cumprod[t.long()].view(-1, 1, 1, 1)
        return (x - at.sqrt() * x0b) / (1 - at).sqrt()

    s.set_timesteps(50)
    outb = sample_loop(model_b, s, 50, xT)
    assert float(outb.max()) < 1.7 - 0.1


def test_ddpm_known_answers():
    """DDPMOracle against closed forms of Ho et al. 2020: posterior-mean coefficients sum rule, beta-tilde variance,
    the noise-free last step, and recovery of a point mass when the ancestral noise is switched off."""
    from oracle.schedulers import DDPMOracle

    s = DDPMOracle(1000, 1e-4, 0.02)
    assert s.timesteps[:3].tolist() == [999, 998, 997] and s.num_inference_steps == 1000
    betas = np.linspace(1e-4, 0.02, 1000)
    ac = np.cumprod(1 - betas)
    # one full-resolution step at t = 500 with zero noise: x_prev = c_x0 * x0_hat + c_xt * x_t, closed-form coefficients
    t = 500
    c_x0 = np.sqrt(ac[t - 1]) * betas[t] / (1 - ac[t])
    c_xt = np.sqrt(1 - betas[t]) * (1 - ac[t - 1]) / (1 - ac[t])
    x = torch.tensor([[[[0.3, -0.2]]]])
    e = torch.tensor([[[[0.5, 1.5]]]])
    x0_hat = ((x - np.sqrt(1 - ac[t]) * e) / np.sqrt(ac[t])).clamp(-1, 1)
    out = s.step(e, t, x, noise=torch.zeros_like(x)).prev_sample
    assert torch.allclose(out, c_x0 * x0_hat + c_xt * x, rtol=2e-5, atol=1e-6)
    # variance: beta-tilde_t = (1 - abar_{t-1}) / (1 - abar_t) * beta_t, applied as sigma * noise
    z = torch.ones_like(x)
    out_z = s.step(e, t, x, noise=z).prev_sample
    sigma = np.sqrt((1 - ac[t - 1]) / (1 - ac[t]) * betas[t])
    assert torch.allclose(out_z - out, torch.full_like(x, float(sigma)), rtol=1e-4)
    # t = 0: no noise is added and the step returns the clipped x0 estimate
    out0 = s.step(e, 0, x, noise=torch.full_like(x, 1e3)).prev_sample
    x0_0 = ((x - np.sqrt(1 - ac[0]) * e) / np.sqrt(ac[0])).clamp(-1, 1)
    assert torch.allclose(out0, x0_0, rtol=1e-5, atol=1e-6)
    # strided sampling (50 steps, "leading"): exact epsilon + zero ancestral noise recovers the point mass
    s.set_timesteps(50)
    assert s.timesteps[:3].tolist() == [980, 960, 940] and int(s.timesteps[-1]) == 0
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 0.9]]]])
    g = torch.Generator().manual_seed(2)
    a = s.alphas_cumprod[980]
    xt = a.sqrt() * x0 + (1 - a).sqrt() * torch.randn(x0.shape, generator=g)
    for tt in s.timesteps:
        at = s.alphas_cumprod[int(tt)]
        eps = (xt - at.sqrt() * x0) / (1 - at).sqrt()
        xt = s.step(eps, tt, xt, noise=torch.zeros_like(xt)).prev_sample
    assert float((xt - x0).abs().max()) < 2e-5


def test_ddim_add_noise():
    s = DDIMOracle(1000)
    x0 = torch.ones(2, 1, 2, 2)
    n = torch.full((2, 1, 2, 2), 2.0)
    t = torch.tensor([0, 999])
    out = s.add_noise(x0, n, t)
    for b in range(2):
        a = s.alphas_cumprod[t[b]]
        assert torch.allclose(out[b], a.sqrt() + 2 * (1 - a).sqrt())


def test_dpmpp_schedule_and_point_mass():
    s = DPMSolverPPOracle(1000, 1e-4, 0.02)
    s.set_timesteps(20)
    assert s.timesteps.tolist()[:3] == [999, 949, 899] and int(s.timesteps[-1]) == 50
    assert s.sigmas.shape == (21,) and float(s.sigmas[-1]) == 0.0 and s.sigmas.dtype == torch.float32
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 0.9]]]])
    g = torch.Generator().manual_seed(2)
    sig0 = s.sigmas[0]
    al0 = 1 / (sig0 ** 2 + 1).sqrt()
    xT = al0 * x0 + sig0 * al0 * torch.randn(x0.shape, generator=g)
    sig_of_t = {int(t): s.sigmas[i] for i, t in enumerate(s.timesteps)}

    def model(x, t):
        sg = sig_of_t[int(t[0])]
        al = 1 / (sg ** 2 + 1).sqrt()
        return (x - al * x0) / (sg * al)

    out = sample_loop(model, s, 20, xT)
    assert float((out - x0).abs().max()) < 1e-5
    assert s.lower_order_nums == 2


def test_dpmpp_orders():
    s = DPMSolverPPOracle(1000)
    s.set_timesteps(5)
    x = torch.zeros(1, 1, 1, 1)
    calls = []
    first, second = s._first_order, s._second_order
    s._first_order = lambda m0, sample: (calls.append(1), first(m0, sample))[1]
    s._second_order = lambda sample: (calls.append(2), second(sample))[1]
    for t in s.timesteps:
        x = s.step(torch.ones_like(x), t, x).prev_sample
    assert calls == [1, 2, 2, 2, 1]


def _run_point_mass(s, n, x0, seed=3):
    """Exact epsilon of a point-mass data distribution along the scheduler's own sigmas; returns (final x, c) with
    c = the constant noise direction (x_t = alpha_t x0 + sigma_t c for every exact solver)."""
    s.set_timesteps(n)
    g = torch.Generator().manual_seed(seed)
    a0, s0 = s._alpha_sigma(s.sigmas[0])
    c = torch.randn(x0.shape, generator=g)
    x = a0 * x0 + s0 * c
    for i, t in enumerate(s.timesteps):
        a, sg = s._alpha_sigma(s.sigmas[i])
        x = s.step((x - a * x0) / sg, t, x).prev_sample
    return x, c


def test_dpmsolver_noise_prediction_known_answers():
    """algorithm_type "dpmsolver" (the dpmsolver1 / dpmsolver2 aliases): construction with the zero final sigma raises
    as diffusers does; with final_sigmas_type="sigma_min" the solver is exact on a point mass - epsilon is constant
    along the trajectory, so x_end = alpha_min x0 + sigma_min c - for order 1 and order 2, and the order sequence follows
    lower_order_final only below 15 steps."""
    with pytest.raises(ValueError, match="sigma_min"):
        make_scheduler("dpmsolver2")
    with pytest.raises(ValueError, match="sigma_min"):
        DPMSolverPPOracle(1000, algorithm_type="dpmsolver", solver_order=1)
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 0.9]]]])
    for name in ("dpmsolver1", "dpmsolver2"):
        for n in (5, 20):
            s = make_scheduler(name, 1000, {"beta_start": 1e-4, "beta_end": 0.02, "final_sigmas_type": "sigma_min"})
            out, c = _run_point_mass(s, n, x0)
            a_end, s_end = s._alpha_sigma(s.sigmas[-1])
            assert float(s.sigmas[-1]) == pytest.approx(float(((1 - s.alphas_cumprod[0]) / s.alphas_cumprod[0]) ** 0.5))
            assert float((out - (a_end * x0 + s_end * c)).abs().max()) < 2e-5, (name, n)
    s = make_scheduler("dpmsolver2", 1000, {"final_sigmas_type": "sigma_min"})
    for n, want in ((5, [1, 2, 2, 2, 1]), (20, [1] + [2] * 19)):
        s.set_timesteps(n)
        calls = []
        first, second = s._first_order, s._second_order
        s._first_order = lambda m0, sample: (calls.append(1), first(m0, sample))[1]
        s._second_order = lambda sample: (calls.append(2), second(sample))[1]
        x = torch.zeros(1, 1, 1, 1)
        for t in s.timesteps:
            x = s.step(torch.ones_like(x), t, x).prev_sample
        assert calls == want
        s._first_order, s._second_order = first, second


def test_unipc_known_answers():
    """UniPC (bh2, data prediction): same linspace schedule as DPM-Solver++, exact recovery of a point mass (the data
    prediction is constant, so predictor and corrector both return alpha x0 + sigma c and the last step returns x0),
    order sequence [1, 2, ..., 2, 1] with the corrector one order behind, and a closed-form first step: with order 1
    UniP is DDIM in lambda space, x_1 = (sigma_1/sigma_0) x - alpha_1 expm1(-h) m0."""
    from oracle.schedulers import UniPCOracle

    s = UniPCOracle(1000, 1e-4, 0.02)
    s.set_timesteps(20)
    d = DPMSolverPPOracle(1000, 1e-4, 0.02)
    d.set_timesteps(20)
    assert torch.equal(s.timesteps, d.timesteps) and torch.equal(s.sigmas, d.sigmas)
    x0 = torch.tensor([[[[0.25, -0.5], [0.75, 0.9]]]])
    for order in (1, 2):
        for n in (2, 5, 20):
            out, _ = _run_point_mass(UniPCOracle(1000, 1e-4, 0.02, solver_order=order), n, x0)
            assert float((out - x0).abs().max()) < 1e-5, (order, n)
    # orders used by predictor / corrector over a 5-step run
    s = UniPCOracle(1000)
    s.set_timesteps(5)
    pred, corr = [], []
    p0, c0 = s._predict, s._correct
    s._predict = lambda sample, order: (pred.append(order), p0(sample, order))[1]
    s._correct = lambda m, last, order: (corr.append(order), c0(m, last, order))[1]
    x = torch.full((1, 1, 1, 1), 0.3)
    g = torch.Generator().manual_seed(0)
    for t in s.timesteps:
        x = s.step(torch.randn(1, 1, 1, 1, generator=g), t, x).prev_sample
    assert pred == [1, 2, 2, 2, 1] and corr == [1, 2, 2, 2]
    assert torch.isfinite(x).all()
    # first step in closed form
    s = UniPCOracle(1000, 1e-4, 0.02)
    s.set_timesteps(10)
    x = torch.tensor([[[[0.4, -1.2]]]])
    e = torch.tensor([[[[0.7, 0.1]]]])
    a0, s0 = s._alpha_sigma(s.sigmas[0].double())
    a1, s1 = s._alpha_sigma(s.sigmas[1].double())
    h = (a1.log() - s1.log()) - (a0.log() - s0.log())
    m0 = (x.double() - s0 * e.double()) / a0
    want = (s1 / s0) * x.double() - a1 * torch.expm1(-h) * m0
    got = s.step(e, s.timesteps[0], x).prev_sample
    assert torch.allclose(got.double(), want, rtol=1e-5, atol=1e-6)
    # the corrector is consistent: fed the exact data prediction at the new point it leaves an exact sample unchanged
    # (covered by the point-mass runs above) and it changes an inexact one
    s = UniPCOracle(1000, 1e-4, 0.02)
    s.set_timesteps(10)
    x = torch.tensor([[[[0.4, -1.2]]]])
    x1 = s.step(e, s.timesteps[0], x).prev_sample
    x2 = s.step(e * 0.5, s.timesteps[1], x1).prev_sample
    assert not torch.equal(s.last_sample, x1) and torch.isfinite(x2).all()


def test_select_timesteps():
    s = make_scheduler("ddim")
    s.set_timesteps(50)
    sel = select_timesteps(s.timesteps, start_step=700)
    assert int(sel[0]) == 700 and int(sel[-1]) == 0
    assert select_timesteps(s.timesteps, last_n_steps=3).tolist() == [40, 20, 0]
    with pytest.raises(ValueError):
        select_timesteps(s.timesteps, start_step=-1)
    with pytest.raises(ValueError):
        select_timesteps(s.timesteps[s.timesteps < 0])
