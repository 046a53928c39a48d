"""Parity of the B200 module mirror against the oracle (fp32 restatement of the reference, oracle/denoiser.py):
blocks, full denoisers, scheduler steps (bit-exact) and the graph-replayed sampling loop."""
import json
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import denoiser as OD  # noqa: E402
from oracle.sampling import make_scheduler, sample_loop  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

MNIST_UNET = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
              "block_out_channels": [64, 128, 128],
              "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
              "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
LDCT_SMALL = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
              "block_out_channels": [128, 128, 256, 256, 512, 512],
              "down_block_types": ["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"],
              "up_block_types": ["UpBlock2D", "AttnUpBlock2D"] + ["UpBlock2D"] * 4}
COMPVIS_SMALL = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2],
                 "model_channels": 64, "attention_resolutions": [], "block_out_channels": [64, 64, 128, 128]}
# configs/LDCT/LDCT_flow_matching_compvis.json (EfficientUNetND, scale-shift ResBlocks, SpatialSelfAttention middle block)
COMPVIS_LDCT = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2, 4, 4],
                "model_channels": 128, "attention_resolutions": [], "block_out_channels": [128, 128, 256, 256, 512, 512]}
TOL_STEP = 1e-2     # north star: per-step velocity / epsilon within 1e-2 relative L2 (bf16 path vs fp32 oracle)
PSNR_MIN = 40.0     # north star: final samples >= 40 dB PSNR


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def psnr(a, b, peak=1.0):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


def build(cfg, conditioning, seed=1):
    from fmdm_b200.models.generators import DiffusionUNetFactory

    model = DiffusionUNetFactory().build(cfg, conditioning, 1)
    sd = OD.reinit_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    return model, {k: v.to(DEV) for k, v in sd.items()}


def test_resblock_variants():
    from fmdm_b200.nn import ResBlockND

    g = torch.Generator().manual_seed(0)
    for (cin, cout, kw) in [(128, 128, dict(emb_activation_before_proj=True, add_embedding_to_hidden=True)),
                            (64, 128, dict(emb_activation_before_proj=True, add_embedding_to_hidden=True)),
                            (128, 64, dict(use_scale_shift_norm=True)),
                            (128, 128, dict(use_scale_shift_norm=True, use_conv=True)),
                            (64, 128, dict(use_conv=True))]:
        blk = ResBlockND(cin, 256, 0.0, out_channels=cout, zero_init_last_conv=False, **kw)
        sd = OD.reinit_state_dict(blk.state_dict(), 3)
        blk.load_state_dict(sd)
        blk = blk.to(DEV).eval()
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        x = torch.randn(2, cin, 20, 24, generator=g).to(DEV)
        emb = torch.randn(2, 256, generator=g).to(DEV)
        ref = OD.resblock(sdd, "", x, emb, scale_shift=kw.get("use_scale_shift_norm", False),
                          act_before_proj=kw.get("emb_activation_before_proj", False),
                          add_to_hidden=kw.get("add_embedding_to_hidden", False))
        with torch.no_grad():
            out = blk(x, emb)
        assert out.shape == ref.shape
        assert rel_l2(out, ref) < 8e-3, (cin, cout, kw, rel_l2(out, ref))
    # virtual concat input == concatenated input
    blk = ResBlockND(192, 256, 0.0, out_channels=128, zero_init_last_conv=False, emb_activation_before_proj=True,
                     add_embedding_to_hidden=True)
    sd = OD.reinit_state_dict(blk.state_dict(), 4)
    blk.load_state_dict(sd)
    blk = blk.to(DEV).eval()
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    a = torch.randn(2, 128, 16, 16, generator=g).to(DEV)
    b = torch.randn(2, 64, 16, 16, generator=g).to(DEV)
    emb = torch.randn(2, 256, generator=g).to(DEV)
    ref = OD.resblock(sdd, "", torch.cat([a, b], 1), emb, act_before_proj=True, add_to_hidden=True)
    with torch.no_grad():
        out = blk((a, b), emb)
    assert rel_l2(out, ref) < 8e-3


def test_attention_blocks():
    from fmdm_b200.nn import DiffusersAttentionND, SpatialSelfAttention

    g = torch.Generator().manual_seed(1)
    att = DiffusersAttentionND(128, heads=16)
    sd = OD.reinit_state_dict(att.state_dict(), 5)
    att.load_state_dict(sd)
    att = att.to(DEV).eval()
    x = torch.randn(2, 128, 16, 16, generator=g).to(DEV)
    ref = OD.diffusers_attention({k: v.to(DEV) for k, v in sd.items()}, "", x, 16)
    with torch.no_grad():
        out = att(x)
    assert rel_l2(out, ref) < 8e-3
    ssa = SpatialSelfAttention(128, heads=4, dim_head=64)
    sd = OD.reinit_state_dict(ssa.state_dict(), 6)
    ssa.load_state_dict(sd)
    ssa = ssa.to(DEV).eval()
    ref = OD.spatial_self_attention({k: v.to(DEV) for k, v in sd.items()}, "", x, 4)
    with torch.no_grad():
        out = ssa(x)
    assert rel_l2(out, ref) < 8e-3


@pytest.mark.parametrize("name,cfg,cond,hw,B", [
    ("mnist28_uncond", MNIST_UNET, None, 28, 4),
    ("mnist32_concat", MNIST_UNET, "concatenate", 32, 3),
    ("ldct64_concat", LDCT_SMALL, "concatenate", 64, 2),
    ("ldct160_concat", LDCT_SMALL, "concatenate", 160, 1),  # rows >= 65 px: rolling-row convs with the fused GroupNorm
    ("ldct512_concat", LDCT_SMALL, "concatenate", 512, 2),  # the headline resolution (BASELINE configs[1]), full arch
    # the exact bench configuration: B = 16 per GPU (the strip schedule, hence the fp32 summation order of the fused
    # GroupNorm statistics, depends on the batch size); the fp32 oracle runs on the same GPU (TF32 off)
    ("ldct512_concat_b16", LDCT_SMALL, "concatenate", 512, 16),
    ("compvis32_concat", COMPVIS_SMALL, "concatenate", 32, 2),
    # EfficientUNetND on rows >= 65 px: rolling-row convs with the scale-shift folded into the operand-transform table
    ("compvis_ldct128_concat", COMPVIS_LDCT, "concatenate", 128, 2),
    ("compvis_ldct256_concat", COMPVIS_LDCT, "concatenate", 256, 1),
])
def test_denoiser_forward_parity(name, cfg, cond, hw, B):
    """per-step prediction within 1e-2 relative L2 of the fp32 oracle (north-star tolerance)."""
    model, sd = build(cfg, cond)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
    c = torch.rand(B, 1, hw, hw, generator=g).to(DEV) if cond else None
    for tval in (999.0, 500.5, 1.0):
        t = torch.full((B,), tval, device=DEV)
        with torch.no_grad():
            ref = OD.denoiser_forward(sd, cfg, x, t, conditioning=cond, channels=1, context=c)
            out = model(x, t, context=c)
            out2 = model(torch.cat([x, c], 1), t) if cond else out
        assert out.dtype == torch.float32 and out.shape == ref.shape
        err = rel_l2(out, ref)
        # north-star tolerance, nothing relaxed.  Measured: LDCT arch 0.7-0.97e-2 (0.78e-2 at B=16, 512^2), EfficientUNetND
        # 0.58-0.80e-2 at 128..512 px; the narrow MNIST arch 0.66-0.88e-2 with its split-bf16 weights (it sat at
        # 0.98-1.00e-2, the bf16 noise floor, with plain bf16 weights: `BaseUNetND.set_weight_split`).
        assert err < TOL_STEP, (name, tval, err)
        worst_sample = max(rel_l2(out[i], ref[i]) for i in range(B))
        assert worst_sample < TOL_STEP, (name, tval, worst_sample)
        assert torch.equal(out, out2)


def test_scheduler_steps_bit_exact():
    """K4: every scheduler step matches the oracle bit for bit in fp32 (north star: within 1 ulp)."""
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override

    g = torch.Generator().manual_seed(8)
    shape = (3, 1, 33, 31)
    for name, n in (("flowmatch", 50), ("ddim", 50), ("dpmsolver++", 20), ("dpmsolver++", 5), ("dpmsolver1", 20),
                    ("dpmsolver2", 20), ("dpmsolver2", 5), ("unipc", 20), ("unipc", 5), ("unipc", 2)):
        ov = resolve_scheduler_override(name)
        params = {"beta_start": 1e-4, "beta_end": 0.02}
        params.update(ov.get("params", {}))
        mine, _ = build_scheduler({"name": ov["name"], "params": params, "num_train_timesteps": 1000}, {})
        orc = make_scheduler(name, 1000, params)
        mine.set_timesteps(n)
        orc.set_timesteps(n)
        assert torch.equal(mine.timesteps, orc.timesteps)
        x_cpu = torch.randn(shape, generator=g) * 1.5
        x_gpu = x_cpu.to(DEV)
        for t in orc.timesteps:
            pred = torch.randn(shape, generator=g)
            x_cpu = orc.step(pred, t, x_cpu).prev_sample
            x_gpu = mine.step(pred.to(DEV), t, x_gpu).prev_sample
            assert torch.equal(x_gpu.cpu(), x_cpu), (name, n, float(t), float((x_gpu.cpu() - x_cpu).abs().max()))
    # ddpm (ancestral): the step's Gaussian draw is passed to both sides
    for n in (1000, 50):
        mine, _ = build_scheduler({"name": "ddpm", "params": {"beta_start": 1e-4, "beta_end": 0.02}}, {})
        orc = make_scheduler("ddpm", 1000, {"beta_start": 1e-4, "beta_end": 0.02})
        mine.set_timesteps(n)
        orc.set_timesteps(n)
        assert torch.equal(mine.timesteps, orc.timesteps)
        x_cpu = torch.randn(shape, generator=g) * 1.5
        x_gpu = x_cpu.to(DEV)
        for t in list(orc.timesteps[:6]) + list(orc.timesteps[-6:]):
            pred = torch.randn(shape, generator=g)
            nz = torch.randn(shape, generator=g)
            x_cpu = orc.step(pred, t, x_cpu, noise=nz).prev_sample
            x_gpu = mine.step(pred.to(DEV), t, x_gpu, noise=nz.to(DEV)).prev_sample
            assert torch.equal(x_gpu.cpu(), x_cpu), ("ddpm", n, float(t), float((x_gpu.cpu() - x_cpu).abs().max()))
    # add_noise
    mine, _ = build_scheduler({"name": "ddim", "params": {"beta_start": 1e-4, "beta_end": 0.02}}, {})
    orc = make_scheduler("ddim", 1000, {"beta_start": 1e-4, "beta_end": 0.02})
    x0 = torch.randn(shape, generator=g)
    nz = torch.randn(shape, generator=g)
    ts = torch.tensor([0, 500, 999])
    assert torch.equal(mine.add_noise(x0.to(DEV), nz.to(DEV), ts).cpu(), orc.add_noise(x0, nz, ts))


@pytest.mark.parametrize("sched,steps", [("flowmatch", 50), ("ddim", 50), ("dpmsolver++", 20), ("dpmsolver2", 20),
                                         ("unipc", 20)])
def test_sampling_loop_parity(sched, steps):
    """Graph-replayed sampling vs the oracle loop (fp32 oracle denoiser).

    flowmatch (the headline sampler): final samples >= 40 dB PSNR against the oracle run.
    ddim / dpmsolver++: with random-init epsilon weights x0 = (x - sqrt(1-a) eps)/sqrt(a) is amplified ~130x and
    clamped, so final samples are sign patterns and a final-sample PSNR is ill-conditioned; instead the B200 path is
    teacher-forced along the ORACLE trajectory: at every step its prediction is within 1e-2 relative L2 of the
    oracle's and its scheduler update of the oracle's prediction is bit-identical.
    All three: the CUDA-graph path and the step-by-step path of the product give bit-identical samples."""
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override, sample_with_scheduler

    model, sd = build(MNIST_UNET, "concatenate", seed=2)
    g = torch.Generator().manual_seed(9)
    B, hw = 4, 32
    noise = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
    cond = torch.rand(B, 1, hw, hw, generator=g).to(DEV)
    ov = resolve_scheduler_override(sched)
    params = {"beta_start": 1e-4, "beta_end": 0.02}
    params.update(ov.get("params", {}))
    mine, _ = build_scheduler({"name": ov["name"], "params": params}, {})
    timing = {}
    out = sample_with_scheduler(model, mine, steps, noise.shape, torch.device(DEV), conditioning_mode="concatenate",
                                conditioning_batch=cond, init_sample=noise, timing=timing)
    assert timing["model_calls"] == steps and timing["model_seconds"] > 0
    out2 = sample_with_scheduler(model, mine, steps, noise.shape, torch.device(DEV), conditioning_mode="concatenate",
                                 conditioning_batch=cond, init_sample=noise, use_cuda_graph=False)
    assert torch.equal(out, out2), float((out - out2).abs().max())

    orc = make_scheduler(sched, 1000, params)
    mine2, _ = build_scheduler({"name": ov["name"], "params": params}, {})
    mine2.set_timesteps(steps)
    orc.set_timesteps(steps)
    x = noise.cpu()
    worst = 0.0
    for t in orc.timesteps:
        tt = t.expand(B).to(DEV)
        ref_pred = OD.denoiser_forward(sd, MNIST_UNET, x.to(DEV), tt.float(), conditioning="concatenate", channels=1,
                                       context=cond)
        with torch.no_grad():
            my_pred = model(x.to(DEV), tt, context=cond)
        worst = max(worst, rel_l2(my_pred, ref_pred))
        x_next = orc.step(ref_pred.cpu(), t, x).prev_sample
        mine_next = mine2.step(ref_pred, t, x.to(DEV)).prev_sample
        assert torch.equal(mine_next.cpu(), x_next)
        x = x_next
    assert worst < TOL_STEP, (sched, worst)  # measured 0.59 / 0.89 / 0.51e-2 (flowmatch / ddim / dpmsolver++)
    if sched == "flowmatch":
        p = psnr(out.clamp(0, 1).cpu(), x.clamp(0, 1))
        assert p >= PSNR_MIN, (sched, p)
    # start_step / last_n_steps subsets run and agree between the two product paths
    kw = dict(last_n_steps=3) if sched == "flowmatch" else dict(start_step=500)
    a = sample_with_scheduler(model, mine, steps, noise.shape, torch.device(DEV), conditioning_mode="concatenate",
                              conditioning_batch=cond, init_sample=noise, **kw)
    b = sample_with_scheduler(model, mine, steps, noise.shape, torch.device(DEV), conditioning_mode="concatenate",
                              conditioning_batch=cond, init_sample=noise, use_cuda_graph=False, **kw)
    assert torch.isfinite(a).all() and torch.equal(a, b)


def test_full_size_properties():
    """Size-independent properties at the headline size (LDCT arch, 512x512): run-to-run determinism (bit for bit),
    per-sample independence of the batch (sample i of a batch of 3 == the same sample run alone up to the bf16 rounding
    noise: the strip height of the conv schedule, hence the fp32 summation order of the GroupNorm partials, depends on
    the batch size), and graph-replayed == step-by-step sampling for a short trajectory (bit for bit)."""
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    model, _ = build(LDCT_SMALL, "concatenate", seed=4)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(3, 1, 512, 512, generator=g).to(DEV)
    c = torch.rand(3, 1, 512, 512, generator=g).to(DEV)
    t = torch.full((3,), 640.25, device=DEV)
    with torch.no_grad():
        full = model(x, t, context=c)
        again = model(x, t, context=c)
        one = model(x[1:2], t[1:2], context=c[1:2])
    assert torch.equal(full, again)
    assert rel_l2(full[1:2], one) < 1e-2
    assert not torch.equal(full[0:1], full[1:2])
    assert torch.isfinite(full).all() and float(full.std()) > 0
    sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    with torch.no_grad():
        a = sample_with_scheduler(model, sched, 3, tuple(x.shape), torch.device(DEV), conditioning_mode="concatenate",
                                  conditioning_batch=c, init_sample=x)
        b = sample_with_scheduler(model, sched, 3, tuple(x.shape), torch.device(DEV), conditioning_mode="concatenate",
                                  conditioning_batch=c, init_sample=x, use_cuda_graph=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("hw,B", [(256, 1), (512, 1), (512, 16)])
def test_ldct_flowmatch_final_sample_psnr(hw, B):
    """North-star tolerance on the headline architecture: final samples of the full 50-Euler-step flow-matching run
    (full LDCT UNetDiffusersND at 256x256 and at the headline 512x512, graph-replayed B200 path) >= 40 dB PSNR against
    the fp32 oracle loop - including the exact bench configuration, B = 16 at 512x512 (BASELINE configs[1]; every
    sample of the batch is held to the bar, measured 58 dB)."""
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    model, sd = build(LDCT_SMALL, "concatenate", seed=6)
    g = torch.Generator().manual_seed(31)
    steps = 50
    noise = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
    cond = torch.rand(B, 1, hw, hw, generator=g).to(DEV)
    sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    with torch.no_grad():
        out = sample_with_scheduler(model, sched, steps, noise.shape, torch.device(DEV),
                                    conditioning_mode="concatenate", conditioning_batch=cond, init_sample=noise)
    orc = make_scheduler("flowmatch", 1000, {})
    orc.set_timesteps(steps)
    x = noise.clone()
    with torch.no_grad():
        for t in orc.timesteps:
            pred = OD.denoiser_forward(sd, LDCT_SMALL, x, t.expand(B).to(DEV).float(), conditioning="concatenate",
                                       channels=1, context=cond)
            x = orc.step(pred, t, x).prev_sample
    for i in range(B):
        p = psnr(out[i].clamp(0, 1), x[i].clamp(0, 1))
        assert p >= PSNR_MIN, (hw, B, i, p)


def test_config3_ddim_and_dpmsolver_final_sample_psnr():
    """BASELINE configs[3]: the DDPM-trained LDCT UNet (`configs/LDCT/LDCT_ddpm_diffusers_nd.json`: full architecture,
    256x256, betas 1e-4..0.02) sampled with `--scheduler ddim` (50 steps) and `--scheduler dpmsolver++` (20 steps):
    final samples of the graph-replayed B200 path >= 40 dB PSNR against the fp32 oracle loop on the same weights, and
    every step's epsilon within 1e-2 relative L2 along the oracle trajectory.  The weights are a conditioned fixture:
    200 seeded epsilon-target training steps on synthetic LDCT pairs (tests/_fixtures.py; measured loss 0.009,
    DDIM 59.5 dB, DPM-Solver++ 48.6 dB, per-step error <= 0.22e-2)."""
    from _fixtures import synthetic_pair, train_epsilon_denoiser
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override, sample_with_scheduler

    hw, B = 256, 2
    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(LDCT_SMALL, "concatenate", 1).to(DEV)
    loss = train_epsilon_denoiser(model, hw=hw, batch=16, steps=200)
    assert loss < 0.05, f"the epsilon fixture did not train (loss {loss})"
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    geval = torch.Generator(device=DEV).manual_seed(77)
    _, cond = synthetic_pair(B, hw, geval)
    noise = torch.randn(B, 1, hw, hw, generator=geval, device=DEV)
    for sname, steps in (("ddim", 50), ("dpmsolver++", 20)):
        ov = resolve_scheduler_override(sname)
        params = {"beta_start": 1e-4, "beta_end": 0.02}
        params.update(ov.get("params", {}))
        mine, _ = build_scheduler({"name": ov["name"], "params": params}, {})
        with torch.no_grad():
            out = sample_with_scheduler(model, mine, steps, noise.shape, torch.device(DEV),
                                        conditioning_mode="concatenate", conditioning_batch=cond, init_sample=noise)
        orc = make_scheduler(sname, 1000, {"beta_start": 1e-4, "beta_end": 0.02})
        orc.set_timesteps(steps)
        x = noise.clone()
        worst = 0.0
        with torch.no_grad():
            for t in orc.timesteps:
                tt = t.expand(B).to(DEV).float()
                pred = OD.denoiser_forward(sd, LDCT_SMALL, x, tt, conditioning="concatenate", channels=1, context=cond)
                worst = max(worst, rel_l2(model(x, tt, context=cond), pred))
                x = orc.step(pred.cpu(), t, x.cpu()).prev_sample.to(DEV)
        assert worst < TOL_STEP, (sname, worst)
        for i in range(B):
            p = psnr(out[i].clamp(0, 1), x[i].clamp(0, 1))
            assert p >= PSNR_MIN, (sname, i, p)


def test_ddpm_graph_sampler_statistics():
    """`--scheduler ddpm` (the reference's default name) through the graph-replayed sampler: a zero-epsilon denoiser makes
    every step an explicit linear-Gaussian map, so the mean and variance of the final sample are known in closed form
    (the mean contracts towards clip(x0_hat) and fresh noise enters every replay); two runs draw different noise."""
    from fmdm_b200.pipelines.utils import GraphSampler, build_scheduler

    class ZeroEps(torch.nn.Module):
        def forward(self, x, t, context=None, t_table=None, step_dev=None):
            return torch.zeros_like(x)

    sched, _ = build_scheduler({"name": "ddpm", "params": {"beta_start": 1e-4, "beta_end": 0.02}}, {})
    sched.set_timesteps(20)
    shape = (4, 1, 64, 64)
    gs = GraphSampler(ZeroEps(), sched, shape, torch.device(DEV))
    init = torch.zeros(shape, device=DEV)
    torch.manual_seed(0)
    a = gs.run(init, None, sched.timesteps)
    b = gs.run(init, None, sched.timesteps)
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    # with eps = 0 and x_T = 0: x0_hat = x / sqrt(abar) (clipped), the chain stays zero-mean; the last step (t = 0,
    # sigma = 0) returns clip(x / sqrt(abar_0)), so the output is bounded by the clip range and not degenerate
    assert abs(float(a.mean())) < 0.05 and float(a.abs().max()) <= 1.0 + 1e-6 and float(a.std()) > 0.05


CA_DIFFUSERS = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
                "block_out_channels": [64, 128], "cross_attention_dim": 4,
                "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D"], "mid_block_type": "UNetMidBlock2DCrossAttn",
                "up_block_types": ["CrossAttnUpBlock2D", "UpBlock2D"]}
CA_EFFICIENT = {"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1,
                "channel_mult": [1, 2], "model_channels": 64, "block_out_channels": [64, 128],
                "attention_resolutions": [2], "cross_attention_resolutions": [2], "cross_attention_in_middle": True,
                "cross_attention_dim": 4, "use_linear_attn": False}


CA_EFFICIENT_LINEAR = {"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1,
                       "channel_mult": [1, 2], "model_channels": 64, "block_out_channels": [64, 128],
                       "attention_resolutions": [1, 2], "cross_attention_resolutions": [2],
                       "cross_attention_in_middle": True, "cross_attention_dim": 4}


def test_linear_attention_kernel():
    """LinearQKVAttention (`attention.py:53-70`) against its fp32 definition, self- and cross-shaped."""
    from fmdm_b200.nn.blocks.attention import LinearQKVAttention

    g = torch.Generator().manual_seed(9)
    att = LinearQKVAttention()
    for b, h, tq, tk, d in ((2, 4, 64, 64, 64), (1, 4, 256, 100, 64), (3, 2, 33, 77, 8), (2, 8, 128, 128, 32),
                            (1, 1, 300, 40, 16)):
        q = torch.randn(b, h, tq, d, generator=g).to(DEV).to(torch.bfloat16)
        k = torch.randn(b, h, tk, d, generator=g).to(DEV).to(torch.bfloat16)
        v = torch.randn(b, h, tk, d, generator=g).to(DEV).to(torch.bfloat16)
        out = att(q, k, v)
        ref = OD._linear_qkv_attention(q.float(), k.float(), v.float())
        assert out.shape == ref.shape and rel_l2(out, ref) < 6e-3, (b, h, tq, tk, d, rel_l2(out, ref))


@pytest.mark.parametrize("name,cfg", [("ca_diffusers_nd", CA_DIFFUSERS), ("ca_efficient_nd", CA_EFFICIENT),
                                      ("ca_efficient_nd_linear", CA_EFFICIENT_LINEAR)])
def test_cross_attention_conditioning_parity(name, cfg):
    """conditioning: "attention" (SURVEY 8f N4): cross-attention blocks over a latent context, against the oracle
    (itself pinned to the reference's golden outputs for the same configs) and against the golden fixture directly."""
    from fmdm_b200.models.generators import DiffusionUNetFactory

    model = DiffusionUNetFactory().build(cfg, "attention", 1)
    sd = OD.reinit_state_dict(model.state_dict(), 13)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(31)
    for hw, ctx_shape in ((32, (2, 4, 8, 8)), (64, (2, 4, 16, 16)), (32, (2, 4, 50)), (32, (2, 50, 4))):
        x = torch.randn(2, 1, hw, hw, generator=g).to(DEV)
        ctx = torch.randn(ctx_shape, generator=g).to(DEV)
        t = torch.tensor([812.0, 33.0], device=DEV)
        ref = OD.denoiser_forward(sdd, cfg, x, t, conditioning="attention", channels=1, context_ca=ctx)
        with torch.no_grad():
            out = model(x, t, context_ca=ctx)
            again = model(x, t, context_ca=ctx)          # keys/values of the same context object come from the cache
        assert out.shape == ref.shape and torch.equal(out, again)
        assert rel_l2(out, ref) < TOL_STEP, (name, hw, ctx_shape, rel_l2(out, ref))
        ctx.mul_(0.5)                                     # in-place change of the context invalidates the cache
        ref2 = OD.denoiser_forward(sdd, cfg, x, t, conditioning="attention", channels=1, context_ca=ctx)
        with torch.no_grad():
            out2 = model(x, t, context_ca=ctx)
        assert rel_l2(out2, ref2) < TOL_STEP and rel_l2(out2, out) > 1e-3
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"denoiser_{name}.pt"),
                      weights_only=False)
    with torch.no_grad():
        out = model(gold["x"].to(DEV), gold["t"].to(DEV), context_ca=gold["context_ca"].to(DEV))
    # measured 0.544 / 0.953 / 0.909e-2 (split-bf16 weights, <= 128 channels; 0.518 / 0.917 / 0.882e-2 when every A tile
    # is fetched once per weight half, FMDM_CONV_NO_DUP=1 - the shared-A form costs ~4 % of the error budget)
    assert rel_l2(out.cpu(), gold["out"]) < TOL_STEP


def test_context_kv_kernel():
    """GroupNorm over the context tokens + tiny-K projection, both output layouts, against torch fp32."""
    from fmdm_b200 import ops

    g = torch.Generator().manual_seed(5)
    for cc, tc, o, groups in ((4, 64, 256, 4), (4, 1000, 1024, 4), (8, 77, 512, 8), (16, 256, 96, 16), (6, 33, 64, 2)):
        ctx = (torch.randn(3, cc, tc, generator=g) * 2 + 0.5).to(DEV)
        gamma, beta = torch.rand(cc, generator=g).to(DEV) + 0.5, torch.randn(cc, generator=g).to(DEV)
        w, bias = (torch.randn(o, cc, generator=g) * 0.3).to(DEV), torch.randn(o, generator=g).to(DEV)
        ref = torch.nn.functional.linear(torch.nn.functional.group_norm(ctx, groups, gamma, beta, 1e-5).transpose(1, 2),
                                         w, bias)                                     # [B][Tc][O]
        tm = ops.context_kv(ctx, gamma, beta, w, bias, groups=groups, eps=1e-5, channel_major=False)
        cm = ops.context_kv(ctx, gamma, beta, w, bias, groups=groups, eps=1e-5, channel_major=True)
        assert tm.shape == (3, tc, o) and cm.shape == (3, o, tc)
        assert rel_l2(tm, ref) < 4e-3 and rel_l2(cm, ref.transpose(1, 2)) < 4e-3


def test_attention_conditioned_sampling_graph_matches_stepwise():
    """conditioning: "attention" through `sample_with_scheduler`: the graph-replayed run (keys/values of the context
    refreshed in place between runs) is bit-identical to the step-by-step run, for two different contexts in a row."""
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    for cfg in (CA_DIFFUSERS, CA_EFFICIENT):
        model = DiffusionUNetFactory().build(cfg, "attention", 1)
        model.load_state_dict(OD.reinit_state_dict(model.state_dict(), 21))
        model = model.to(DEV).eval()
        sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
        g = torch.Generator().manual_seed(3)
        init = torch.randn(2, 1, 32, 32, generator=g).to(DEV)
        for k in range(2):
            ctx = (torch.randn(2, 4, 8, 8, generator=g) * (1.0 + k)).to(DEV)
            a = sample_with_scheduler(model, sched, 6, tuple(init.shape), torch.device(DEV), conditioning_mode="attention",
                                      conditioning_batch=ctx, latent_norm="standardize", init_sample=init)
            b = sample_with_scheduler(model, sched, 6, tuple(init.shape), torch.device(DEV), conditioning_mode="attention",
                                      conditioning_batch=ctx, latent_norm="standardize", init_sample=init,
                                      use_cuda_graph=False)
            assert torch.isfinite(a).all() and torch.equal(a, b), (cfg["unet_impl"], k, float((a - b).abs().max()))
