#!/usr/bin/env python
"""bench.py — LDCT 512x512 conditional flow-matching samples/s @50 Euler steps (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            B200 arm (this repo's kernels)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle port of the reference's path
  (N > 1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`)

One bench "step" = one full 50-Euler-step sampling run of one batch (16 samples per GPU) of synthetic LDCT-shaped
input through `UNetDiffusersND` (configs/LDCT/LDCT_flow_matching_diffusers_nd.json, conditioning "concatenate",
seeded random weights) and the flow-match Euler scheduler.  Prints ONE JSON line on rank 0.

Besides the headline the line carries
  roofline      the dominant kernel timed INSIDE steady state (profiled forwards queued behind ~1.5 s of back-to-back
                sampling steps, so they run at the loop's power-capped clocks) against the sustained peak; the same
                kernel after an idle pause against the burst peak; and the sum of all kernels of a forward against the
                measured time of one Euler step
  cpu_baseline  the oracle port of the path on the box's host cores (bounded sample)
  extra         the other BASELINE configs under the same clock: configs[4] training step (every N; the data-parallel
                gradient all-reduce at N > 1), and at N = 1 configs[0] MNIST, configs[2] latent + KL decode,
                configs[3] DDIM-50 / DPM-Solver++-20, each with its own bounded CPU baseline
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LDCT_UNET = {
    "unet_impl": "diffusers_nd", "sample_size": 256, "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
    "block_out_channels": [128, 128, 256, 256, 512, 512],
    "down_block_types": ["DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
    "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"],
    "attention_resolutions": [], "cross_attention_resolutions": [], "emb_activation_before_proj": False,
}
MNIST_UNET = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
              "block_out_channels": [64, 128, 128], "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
              "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
KL_VAE = {"in_channels": 1, "out_channels": 1, "resolution": 256, "down_channels": [128, 256, 512, 512],
          "num_res_blocks": 2, "z_channels": 4, "embed_dim": 4, "attn_heads": 4, "attn_dim_head": 64}
# SURVEY.md §8d, FlopCounterMode on the reference modules
FLOP_PER_SAMPLE_FWD = 1.9944e12      # LDCT 512^2 (conv 1.9723e12)
FLOP_LDCT256_FWD = 4.9657e11
FLOP_MNIST28_FWD = 2.3969e9
FLOP_LATENT_FWD, FLOP_KL_DECODE = 3.1091e10, 2.4918e12
EULER_STEPS = 50
IMG = 512
BATCH_PER_GPU = 16
METRIC = "LDCT 512^2 flow-matching samples/s @50 Euler steps"
WORKLOAD = ("LDCT 512x512 concat flow-matching UNetDiffusersND (128,128,256,256,512,512), 50 Euler steps "
            "(BASELINE configs[1])")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period: float = 0.2):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        mx = max([float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()] or [0.0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.samples)}


def synthetic_inputs(batch: int, seed: int, device="cpu", hw: int = IMG, channels: int = 1):
    """noise ~ N(0,1); conditioning = clamp(u + 0.05 n, 0, 1), u ~ U[0,1]: LDCT-shaped, in [0,1]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shape = (batch, channels, hw, hw)
    noise = torch.randn(shape, generator=g)
    cond = torch.rand(shape, generator=g) + 0.05 * torch.randn(shape, generator=g)
    return noise.to(device), cond.clamp_(0, 1).to(device)


def manifest_state_dict(name: str, seed: int = 0):
    """Reference-format state_dict from the committed key/shape manifest of the REFERENCE's own module
    (tests/golden/state_keys_*.json, written by oracle/make_golden*.py from /root/reference/src) and the oracle's seeded
    per-key initialiser: both arms use it, so the CPU arm builds its weights without importing this package."""
    from oracle import denoiser as OD

    with open(os.path.join(GOLDEN, f"state_keys_{name}.json")) as f:
        meta = json.load(f)
    return OD.reinit_state_dict({k: torch.zeros(shape) for k, shape in meta["keys"]}, seed), meta


def _cpu_setup():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference modules restated in oracle/denoiser.py + schedulers)
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(n_euler: int, repeats: int, warmup: int, batch: int = 1, *, cfg=None, manifest="ldct_diffusers_nd",
                  hw: int = IMG, sched: str = "flowmatch", full_steps: int = EULER_STEPS, cond: bool = True):
    """samples/s of the CPU path extrapolated from per-step time on a bounded sample (`batch`, last n_euler steps)."""
    from oracle import denoiser as OD
    from oracle.sampling import make_scheduler, sample_loop

    cores = _cpu_setup()
    sd, meta = manifest_state_dict(manifest)
    cfg = cfg or meta["cfg"]
    noise, cnd = synthetic_inputs(batch, 42, hw=hw)
    conditioning = "concatenate" if cond else None

    def model(inp, t):
        return OD.unet_diffusers_nd_forward(sd, cfg, inp[:, :1], t, conditioning=conditioning, channels=1,
                                            context=inp[:, 1:] if cond else None)

    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            sch = make_scheduler(sched, 1000, {"beta_start": 1e-4, "beta_end": 0.02} if sched != "flowmatch" else {})
            t0 = time.perf_counter()
            sample_loop(model, sch, full_steps, noise, cnd if cond else None, last_n_steps=n_euler)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    per_step = sum(times) / len(times) / n_euler
    return batch / (per_step * full_steps), per_step, cores


def gpu_eager_rate(dev, batch: int, mode: str, repeats: int = 3):
    """Reported context, not the product: the oracle's plain-PyTorch restatement of the reference denoiser run EAGERLY
    on the same B200 (cuDNN / cuBLAS / ATen kernels, NCHW fp32 with TF32 convs, or autocast bf16 + channels_last),
    i.e. what the unmodified reference modules would do on this GPU.  samples/s extrapolated from the forward time
    (the scheduler step is < 0.1 % of it)."""
    from oracle import denoiser as OD

    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in manifest_state_dict("ldct_diffusers_nd")[0].items()}
    noise, cond = synthetic_inputs(batch, 42, dev)
    if mode == "bf16":
        sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
        noise = noise.contiguous(memory_format=torch.channels_last)
        cond = cond.contiguous(memory_format=torch.channels_last)
    t = torch.full((batch,), 500.0, device=dev)

    def fwd():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            return OD.unet_diffusers_nd_forward(sd, LDCT_UNET, noise, t, conditioning="concatenate", channels=1,
                                                context=cond)

    with torch.no_grad():
        for _ in range(2):
            fwd()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(repeats):
            fwd()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / repeats
    return batch / (ms * 1e-3 * EULER_STEPS), ms


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_euler = 1
    t_begin = time.perf_counter()
    rate, per_euler, cores = cpu_port_rate(n_euler, repeats=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_euler * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "euler_steps": EULER_STEPS, "sample_batch": 1,
                   "euler_steps_timed_per_bench_step": n_euler,
                   "note": "bounded sample: B=1, one Euler step per bench step, samples/s = 1/(t_euler*50); weights from "
                           "the reference module's key/shape manifest, no module of the B200 package is imported"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"B=1, {n_euler} Euler step(s) of 50 per bench step at 512x512, fp32, torch CPU "
                                   f"({cores} threads); samples/s = 1/(t_euler*50)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_begin,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs under the same clock
# ------------------------------------------------------------------------------------------------------------------
def _timed_runs(fn, dev, reps: int, warm: int = 2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / 1e3 / reps, out


def extra_sampling_config(dev, peaks, *, name, cfg, manifest, cond, B, hw, sched, steps, flop_fwd, cpu, reps=3):
    """One other sampling config through the public API (`sample_with_scheduler`, graph-replayed), inputs resident."""
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override, sample_with_scheduler

    sd, meta = manifest_state_dict(manifest)
    cfg = cfg or meta["cfg"]
    model = DiffusionUNetFactory().build(cfg, cond, 1)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    ov = resolve_scheduler_override(sched)
    params = {} if sched == "flowmatch" else {"beta_start": 1e-4, "beta_end": 0.02}
    params.update(ov.get("params", {}))
    sch, _ = build_scheduler({"name": ov["name"], "params": params}, {})
    noise, cnd = synthetic_inputs(B, 7, dev, hw=hw)
    kw = dict(conditioning_mode=cond, conditioning_batch=cnd if cond else None, init_sample=noise)
    with torch.no_grad():
        dt, out = _timed_runs(lambda: sample_with_scheduler(model, sch, steps, tuple(noise.shape), dev, **kw), dev, reps)
    assert torch.isfinite(out).all()
    rec = {"config": name, "value": B / dt, "unit": "samples/s", "batch": B, "steps": steps, "scheduler": sched,
           "s_per_run": dt, "model_tflops": B / dt * flop_fwd * steps / 1e12,
           "frac_of_sustained_peak": B / dt * flop_fwd * steps / 1e12 / peaks["tf_sustained"]}
    if cpu:
        rate, per_step, cores = cpu_port_rate(1, repeats=1, warmup=0, batch=cpu["batch"], cfg=cfg, manifest=manifest, hw=hw,
                                              sched=sched, full_steps=steps, cond=bool(cond))
        rec["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"B={cpu['batch']}, 1 of {steps} steps at {hw}x{hw}, fp32 oracle port; "
                                         f"samples/s = B/(t_step*{steps})"}
    return rec


def extra_latent_config(dev, peaks, cpu: bool, B: int = 128, lat: int = 64, steps: int = EULER_STEPS, reps: int = 2):
    """configs[2]: latent flow matching on AutoencoderKL f=8 latents (64x64x4, concat conditioning latents -> 8 input
    channels), 50 Euler steps, then `AutoencoderKL.decode` to 512x512."""
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.models.vae import AutoencoderKL
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler
    from oracle import denoiser as OD

    torch.manual_seed(0)
    cfg = dict(LDCT_UNET, in_channels=4, out_channels=4)
    unet = DiffusionUNetFactory().build(cfg, "concatenate", 4)
    usd = OD.reinit_state_dict(unet.state_dict(), 0)
    unet.load_state_dict(usd)
    unet = unet.to(dev).eval()
    vsd, vmeta = manifest_state_dict("vae_ldct_kl")
    vae = AutoencoderKL(**KL_VAE)
    vae.load_state_dict(vsd, strict=False)
    vae = vae.to(dev).eval()
    sch, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    x, c = synthetic_inputs(B, 11, dev, hw=lat, channels=4)
    kw = dict(conditioning_mode="concatenate", conditioning_batch=c, init_sample=x)
    with torch.no_grad():
        t_unet, z = _timed_runs(lambda: sample_with_scheduler(unet, sch, steps, tuple(x.shape), dev, **kw), dev, reps)
        t_dec, img = _timed_runs(lambda: vae.raw_output_to_image(vae.decode(z, denorm=True)), dev, reps)
    assert img.shape == (B, 1, 8 * lat, 8 * lat) and torch.isfinite(img).all()
    flop = FLOP_LATENT_FWD * steps + FLOP_KL_DECODE
    total = t_unet + t_dec
    rec = {"config": "configs[2] latent flow matching 64x64x4 (50 Euler) + AutoencoderKL decode to 512x512, batch 128",
           "value": B / total, "unit": "samples/s", "batch": B, "s_unet_50_steps": t_unet, "s_kl_decode": t_dec,
           "model_tflops": B * flop / total / 1e12, "frac_of_sustained_peak": B * flop / total / 1e12 / peaks["tf_sustained"]}
    if cpu:
        from oracle import vae_decoder as OV

        cores = _cpu_setup()
        xc, cc = synthetic_inputs(1, 11, hw=lat, channels=4)
        with torch.no_grad():
            t0 = time.perf_counter()
            OD.unet_diffusers_nd_forward(usd, cfg, xc, torch.full((1,), 500.0), conditioning="concatenate", channels=4,
                                         context=cc)
            t_step = time.perf_counter() - t0
            t0 = time.perf_counter()
            OV.kl_decode(vsd, vmeta["cfg"], xc, denorm=True)
            t_d = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 1.0 / (t_step * steps + t_d), "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"B=1: one latent denoiser forward ({t_step:.2f} s, x{steps}) + one KL decode "
                                         f"({t_d:.2f} s), fp32 oracle port"}
    return rec


def extra_train_step(dev, peaks, rank, world, steps: int, cpu: bool, B: int = 16, hw: int = 256):
    """configs[4]: LDCT 256x256 flow-matching training step (fwd + bwd + AdamW; data-parallel gradient all-reduce),
    batch 16 per GPU, batches taken from pinned host memory and the loss read back every step."""
    import torch.distributed as dist

    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.training import FlowMatchingTrainer

    sd, _ = manifest_state_dict("ldct_diffusers_nd")
    model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    tr = FlowMatchingTrainer(model, lr=1e-4)
    g = torch.Generator().manual_seed(1 + rank)
    h_clean = torch.rand(B, 1, hw, hw, generator=g).pin_memory()
    h_ldct = torch.rand(B, 1, hw, hw, generator=g).pin_memory()
    d_clean, d_ldct = h_clean.to(dev), h_ldct.to(dev)
    h_loss = torch.zeros((), dtype=torch.float32).pin_memory()

    def one():
        d_clean.copy_(h_clean, non_blocking=True)
        d_ldct.copy_(h_ldct, non_blocking=True)
        h_loss.copy_(tr.step(d_clean, d_ldct), non_blocking=True)

    first = None
    for i in range(5):  # 2 eager steps, capture, 2 replays
        one()
        if i == 0:
            torch.cuda.synchronize(dev)
            first = float(h_loss)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    last = float(h_loss)
    # the gradient all-reduce on its own (what an un-overlapped collective would add to every step)
    ar_ms = None
    if world > 1:
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tr.reducer.reduce_all()
        torch.cuda.synchronize(dev)
        f0.record()
        for _ in range(3):
            tr.reducer.reduce_all()
        f1.record()
        torch.cuda.synchronize(dev)
        ar_ms = f0.elapsed_time(f1) / 3
    value = B * world / (ms / 1e3)
    tflops = value * 3 * FLOP_LDCT256_FWD / 1e12 / world
    rec = {"config": "configs[4] LDCT 256x256 flow-matching training step (fwd+bwd+AdamW, gradient all-reduce), "
                     f"batch {B}/GPU", "value": value, "unit": "samples/s", "ms_per_step": ms, "n_ranks": world,
           "steps": steps, "loss_first": first, "loss_last": last, "allreduce_ms_standalone": ar_ms,
           "allreduce_bytes": int(tr.optimizer.flat.numel) * 4, "grad_reduce": getattr(tr, "reduce_mode", "after replay"),
           "model_tflops_per_gpu": tflops, "frac_of_sustained_peak": tflops / peaks["tf_sustained"],
           "e2e": "timed region includes the H2D copy of each batch from pinned memory and the D2H of the loss",
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    tr.reducer.remove()
    if cpu and rank == 0:
        from oracle import training as OT

        cores = _cpu_setup()
        gg = torch.Generator().manual_seed(2)
        batch = (torch.rand(1, 1, hw, hw, generator=gg), torch.rand(1, 1, hw, hw, generator=gg),
                 torch.randn(1, 1, hw, hw, generator=gg), torch.rand(1, generator=gg))
        t0 = time.perf_counter()
        OT.train_steps(sd, LDCT_UNET, [batch], lr=1e-4, weight_decay=0.0)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 1.0 / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"B=1, one fwd+bwd+AdamW step at {hw}x{hw}, fp32 oracle port (torch autograd)"}
    return rec


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def profile_forward(model, noise_d, cond_d, dev, local_rank, gs, peaks, B):
    """Per-kernel CUDA-event timing of eager denoiser forwards (events on the launching stream around every launch).

    steady: the profiled forwards are queued, with no host pause, behind `pre` graph-replayed Euler steps, so their
    kernels run at the clocks of the sampling loop (power-capped); isolated: one forward after the GPU idled (boosted
    clocks) - the pairing the burst peak belongs to."""
    from fmdm_b200 import ops

    t_mid = torch.full((B,), 500.0, device=dev)
    model(noise_d, t_mid, context=cond_d)  # eager path warm (weight packs, attributes)
    torch.cuda.synchronize(dev)
    time.sleep(1.0)
    with ops.profile() as rec_iso:
        model(noise_d, t_mid, context=cond_d)
    pre, reps = 45, 3
    sampler = ClockSampler(local_rank, period=0.1)
    gs.cursor.zero_()
    sampler.start()
    for _ in range(pre):
        gs.graph.replay()
    gs.cursor.zero_()
    with ops.profile() as rec:
        for _ in range(reps):
            model(noise_d, t_mid, context=cond_d)
    clocks = sampler.stop()

    def fold(rows, n):
        by = {}
        for tag, work, ms in rows:
            a = by.setdefault(tag, [0.0, 0.0, 0])
            a[0] += work / n; a[1] += ms / n; a[2] += 1
        for a in by.values():
            a[2] //= n
        return by

    return fold(rec.rows, reps), fold(rec_iso.rows, 1), clocks


def run_b200_arm(args):
    import torch.distributed as dist

    from fmdm_b200 import ops
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.parallel import gather_samples, init_distributed
    from fmdm_b200.pipelines import utils as PU
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    rank, world, local_rank = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the first collective creates the NCCL communicator, and NCCL prints its version banner to STDOUT; keep
        # stdout to the one JSON line by pointing fd 1 at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    peaks = read_peaks()
    B = args.batch
    total = B * world

    sd, _ = manifest_state_dict("ldct_diffusers_nd")
    model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    scheduler, _ = build_scheduler({"name": "flow_match_euler", "num_train_timesteps": 1000, "params": {}}, {})
    noise_h, cond_h = synthetic_inputs(B, 42 + rank)
    noise_h, cond_h = noise_h.pin_memory(), cond_h.pin_memory()
    noise_d, cond_d = noise_h.to(dev), cond_h.to(dev)
    out_h = torch.empty((B, 1, IMG, IMG), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_run_device():
        # inputs already resident in HBM; the run ends with the (N>1) all-gather of the fp32 samples
        x = sample_with_scheduler(model, scheduler, EULER_STEPS, tuple(noise_d.shape), dev,
                                  conditioning_mode="concatenate", conditioning_batch=cond_d, init_sample=noise_d)
        x = ops.clamp_f32(x, 0.0, 1.0)
        return gather_samples(x, total, rank, world)

    def one_run_e2e():
        # public API with HOST buffers: H2D of this step's inputs, sampling, D2H of the samples
        nz = noise_h.to(dev, non_blocking=True)
        cd = cond_h.to(dev, non_blocking=True)
        x = sample_with_scheduler(model, scheduler, EULER_STEPS, tuple(nz.shape), dev,
                                  conditioning_mode="concatenate", conditioning_batch=cd, init_sample=nz)
        x = ops.clamp_f32(x, 0.0, 1.0)
        out_h.copy_(x, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_h

    extra = {}
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            one_run_device()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            one_run_device()
        e1.record()
        barrier()
        clocks = sampler.stop()
        elapsed = e0.elapsed_time(e1) / 1e3

        # e2e through the public API with host buffers
        one_run_e2e()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            one_run_e2e()  # ends with the stream-ordered D2H copy + stream synchronize
        f1.record()
        barrier()
        e2e_elapsed = f0.elapsed_time(f1) / 1e3

        if world > 1:
            t = torch.tensor([elapsed, e2e_elapsed], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed, e2e_elapsed = float(t[0]), float(t[1])

        # graph sampler bookkeeping: kernels launched per Euler step
        gs = max(PU._GRAPH_CACHE.values(), key=lambda g: g.launches_per_step)
        launches_per_euler = gs.launches_per_step
        gpu_launches = launches_per_euler * EULER_STEPS * args.steps + args.steps
        ms_per_step = elapsed / args.steps * 1e3
        ms_per_euler = ms_per_step / EULER_STEPS

        roof = cpu = None
        if rank == 0:
            by, by_iso, prof_clocks = profile_forward(model, noise_d, cond_d, dev, local_rank, gs, peaks, B)
            # dominant kernel = the conv variant with the largest share of the forward (the rolling-row kernel with
            # the fused GroupNorm operand transform on LDCT-512)
            conv_tags = [k for k in by if k.startswith("conv_rolling") or k == "conv_tile"]
            top = max(conv_tags, key=lambda k: by[k][1])
            cw, cms, cn = by[top]
            achieved = cw / (cms * 1e-3) / 1e12
            iw, ims, _ = by_iso[top]
            achieved_iso = iw / (ims * 1e-3) / 1e12
            fwd_ms = sum(v[1] for v in by.values())
            fwd_ms_iso = sum(v[1] for v in by_iso.values())
            allw = sum(by[k][0] for k in conv_tags)
            allms = sum(by[k][1] for k in conv_tags)
            gn = by.get("groupnorm", [0.0, 1e-9, 1])
            traffic = traffic_note = None
            tpath = os.path.join(ROOT, "profiles", "dominant_kernel_dram.json")
            if os.path.exists(tpath):  # dram bytes of one launch of the dominant kernel from the committed ncu capture
                with open(tpath) as f:
                    tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = (f"ncu dram read+write of one launch ({tj.get('instance')}); algorithmic bytes of that "
                                f"launch {tj.get('algorithmic_bytes_per_launch')}, FLOP {tj.get('flop_per_launch')}")
            kernel_names = {"conv_rolling_xf": "conv_rolling_kernel<128,1> (tcgen05 rolling-row implicit GEMM + fused "
                                               "GroupNorm/SiLU operand transform)",
                            "conv_rolling": "conv_rolling_kernel<128,0> (tcgen05 rolling-row implicit GEMM)",
                            "conv_tile": "conv_igemm_persistent_kernel (tcgen05 implicit GEMM, per-tile)"}
            roof = {
                "bound": "tensor", "kernel": kernel_names[top],
                "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["src"] + " sustained",
                "timing": "steady state: CUDA events around every launch of 3 eager forwards queued behind 45 graph-replayed "
                          "Euler steps (no host pause), i.e. at the sampling loop's own clocks",
                "sm_mhz_during_profile": prof_clocks.get("sm_mhz"), "clock_reasons_during_profile": prof_clocks.get("reasons"),
                "flop_per_launch": cw / cn, "avg_launch_ms": cms / cn, "launches_per_forward": cn,
                "share_of_forward": cms / fwd_ms,
                "isolated": {"achieved": achieved_iso, "peak": peaks["tf_burst"], "frac": achieved_iso / peaks["tf_burst"],
                             "avg_launch_ms": ims / cn, "forward_ms": fwd_ms_iso,
                             "note": "same kernel, one forward after the GPU idled 1 s (boosted clocks), against the burst peak"},
                "consistency": {"sum_of_kernels_ms_per_forward": fwd_ms, "ms_per_euler_step_in_run": ms_per_euler,
                                "ratio": fwd_ms / ms_per_euler,
                                "note": "kernels of one profiled forward (steady state) vs 1/50 of the timed run's step; "
                                        "the run adds the scheduler step, the cursor kernel and launch gaps"},
                "all_conv_kernels": {"achieved": allw / (allms * 1e-3) / 1e12, "share_of_forward": allms / fwd_ms,
                                     "frac": allw / (allms * 1e-3) / 1e12 / peaks["tf_sustained"],
                                     "launches_per_forward": sum(by[k][2] for k in conv_tags)},
                "groupnorm": {"bound": "hbm", "achieved": gn[0] / (gn[1] * 1e-3) / 1e9, "peak": peaks["hbm"],
                              "unit": "GB/s", "frac": gn[0] / (gn[1] * 1e-3) / 1e9 / peaks["hbm"],
                              "share_of_forward": gn[1] / fwd_ms,
                              "note": "standalone GroupNorm apply (small levels, attention norms) only: on rows >= 65 "
                                      "px the apply runs inside the consumer conv; algorithmic bytes = 1 read + 1 "
                                      "write bf16"},
                "per_kernel_ms_per_forward": {k: round(v[1], 3) for k, v in by.items()},
                "per_kernel_launches_per_forward": {k: v[2] for k, v in by.items()},
            }
            if args.eager_baseline:
                eager = {}
                for mode in ("fp32_tf32", "bf16"):
                    rate, ms = gpu_eager_rate(dev, B, mode)
                    eager[mode] = {"value": rate, "unit": "samples/s", "forward_ms": ms}
                roof["gpu_eager_reference"] = dict(
                    eager, note="oracle restatement of the reference modules, eager PyTorch on the same B200 "
                                "(fp32_tf32: NCHW fp32 weights, TF32 convs; bf16: autocast + channels_last); context only")
            if not args.no_cpu_baseline:
                rate, per_euler, cores = cpu_port_rate(1, repeats=2, warmup=0)
                cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                       "sample": "B=1, 1 Euler step of 50 at 512x512 (x2 repeats), fp32 torch CPU oracle port, all "
                                 f"host threads ({cores}); extrapolated: samples/s = 1/(t_euler*50)"}

    # ---- extras (after the headline region; never allowed to take the line down) ------------------------------------
    if not args.no_extras:
        PU._GRAPH_CACHE.clear()
        torch.cuda.empty_cache()
        want_cpu = (world == 1) and not args.no_cpu_baseline

        def guarded(key, fn):
            try:
                t0 = time.perf_counter()
                rec = fn()
                if rec is not None:
                    rec["wall_s"] = round(time.perf_counter() - t0, 1)
                    extra[key] = rec
            except Exception as exc:  # noqa: BLE001
                extra[key] = {"error": f"{type(exc).__name__}: {exc}"}
            PU._GRAPH_CACHE.clear()
            torch.cuda.empty_cache()

        guarded("train_step", lambda: extra_train_step(dev, peaks, rank, world, steps=max(args.steps, 10), cpu=want_cpu))
        if world == 1:
            with torch.no_grad():
                guarded("mnist", lambda: extra_sampling_config(
                    dev, peaks, name="configs[0] MNIST 28x28 unconditional flow matching, 50 Euler steps, batch 64",
                    cfg=None, manifest="mnist_diffusers_nd_uncond", cond=None, B=64, hw=28, sched="flowmatch",
                    steps=50, flop_fwd=FLOP_MNIST28_FWD, cpu={"batch": 64} if want_cpu else None, reps=5))
                guarded("latent_kl", lambda: extra_latent_config(dev, peaks, cpu=want_cpu))
                guarded("ddim50_256", lambda: extra_sampling_config(
                    dev, peaks, name="configs[3] LDCT 256x256 DDPM-trained UNet, ddim 50 steps, batch 16",
                    cfg=LDCT_UNET, manifest="ldct_ddpm_diffusers_nd", cond="concatenate", B=16, hw=256, sched="ddim",
                    steps=50, flop_fwd=FLOP_LDCT256_FWD, cpu={"batch": 1} if want_cpu else None))
                guarded("dpmsolverpp20_256", lambda: extra_sampling_config(
                    dev, peaks, name="configs[3] LDCT 256x256 DDPM-trained UNet, dpmsolver++ 20 steps, batch 16",
                    cfg=LDCT_UNET, manifest="ldct_ddpm_diffusers_nd", cond="concatenate", B=16, hw=256,
                    sched="dpmsolver++", steps=20, flop_fwd=FLOP_LDCT256_FWD, cpu=None))

    if rank == 0:
        value = total * args.steps / elapsed
        e2e_value = total * args.steps / e2e_elapsed
        model_tflops = value * FLOP_PER_SAMPLE_FWD * EULER_STEPS / 1e12 / world
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": total, "euler_steps": EULER_STEPS,
                       "l2": "inputs larger than L2 (activations 1-2 GiB per tensor at level 0)",
                       "parallelism": f"batch sharded over {world} GPU(s), final all-gather"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 2 * B * IMG * IMG * 4,
                    "d2h_bytes_per_step": B * IMG * IMG * 4},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "model_tflops_per_gpu": model_tflops,
            "model_frac_of_sustained_peak": model_tflops / peaks["tf_sustained"],
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=("b200", "reference"))
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="samples per GPU (headline: 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (extra.*)")
    ap.add_argument("--eager-baseline", action="store_true",
                    help="also time the reference's plain-PyTorch path eagerly on the GPU (context figure)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
