#!/usr/bin/env python
"""bench.py — LDCT 512x512 conditional flow-matching samples/s @50 Euler steps (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            B200 arm (this repo's kernels)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle port of the reference's path
  (N > 1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`)

One bench "step" = one full 50-Euler-step sampling run of one batch (16 samples per GPU) of synthetic LDCT-shaped
input through `UNetDiffusersND` (configs/LDCT/LDCT_flow_matching_diffusers_nd.json, conditioning "concatenate",
random-init weights, seed 0) and the flow-match Euler scheduler.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
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
FLOP_PER_SAMPLE_FWD = 1.9944e12  # SURVEY.md §8d, FlopCounterMode on the reference module (conv 1.9723e12)
EULER_STEPS = 50
IMG = 512
BATCH_PER_GPU = 16
METRIC = "LDCT 512^2 flow-matching samples/s @50 Euler steps"
WORKLOAD = ("LDCT 512x512 concat flow-matching UNetDiffusersND (128,128,256,256,512,512), 50 Euler steps "
            "(BASELINE configs[1])")


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

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
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
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        mx = max([float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()] or [0.0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.samples)}


def synthetic_inputs(batch: int, seed: int, device="cpu"):
    """noise ~ N(0,1); conditioning = clamp(u + 0.05 n, 0, 1), u ~ U[0,1]: LDCT-shaped, in [0,1]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    noise = torch.randn(batch, 1, IMG, IMG, generator=g)
    cond = (torch.rand(batch, 1, IMG, IMG, generator=g) + 0.05 * torch.randn(batch, 1, IMG, IMG, generator=g))
    return noise.to(device), cond.clamp_(0, 1).to(device)


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference modules restated in oracle/denoiser.py + schedulers)
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(n_euler: int, repeats: int, warmup: int, batch: int = 1):
    """samples/s of the CPU path extrapolated from per-Euler-step time on a bounded sample (B=1, n_euler steps)."""
    from oracle import denoiser as OD
    from oracle.sampling import make_scheduler, sample_loop

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = _random_state_dict()
    noise, cond = synthetic_inputs(batch, 42)

    def model(inp, t):
        return OD.unet_diffusers_nd_forward(sd, LDCT_UNET, inp[:, :1], t, conditioning="concatenate", channels=1,
                                            context=inp[:, 1:])

    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            sch = make_scheduler("flowmatch", 1000)
            t0 = time.perf_counter()
            sample_loop(model, sch, EULER_STEPS, noise, cond, last_n_steps=n_euler)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    per_euler = sum(times) / len(times) / n_euler
    return batch / (per_euler * EULER_STEPS), per_euler, cores


def gpu_eager_rate(dev, batch: int, mode: str, repeats: int = 3):
    """Reported context, not the product: the oracle's plain-PyTorch restatement of the reference denoiser run EAGERLY
    on the same B200 (cuDNN / cuBLAS / ATen kernels, NCHW fp32 with TF32 convs, or autocast bf16 + channels_last),
    i.e. what the unmodified reference modules would do on this GPU.  samples/s extrapolated from the forward time
    (the scheduler step is < 0.1 % of it)."""
    from oracle import denoiser as OD

    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in _random_state_dict().items()}
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


def _random_state_dict():
    """Reference-format random-init state_dict (seed 0) without needing a GPU: built from the module mirror's
    parameter shapes (identical to the reference's under the same seed, tests/test_api_conformance.py)."""
    from fmdm_b200.models.generators import DiffusionUNetFactory

    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1)
    return {k: v.detach() for k, v in model.state_dict().items()}


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
                   "note": "bounded sample: B=1, one Euler step per bench step, samples/s = 1/(t_euler*50)"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"B=1, {n_euler} Euler step(s) of 50 per bench step at 512x512, fp32, torch CPU "
                                   f"({cores} threads); samples/s = 1/(t_euler*50)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_begin,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch.distributed as dist

    from fmdm_b200 import ops
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.parallel import gather_samples, init_distributed
    from fmdm_b200.pipelines.utils import GraphSampler, build_scheduler, sample_with_scheduler

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

    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1).to(dev).eval()
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
        from fmdm_b200.pipelines import utils as PU
        launches_per_euler = max((gs.launches_per_step for gs in PU._GRAPH_CACHE.values()), default=0)
        gpu_launches = launches_per_euler * EULER_STEPS * args.steps + args.steps

        roof = cpu = None
        if rank == 0:
            # live per-kernel timing of ONE eager denoiser forward (CUDA events around every launch on the
            # launching stream); the dominant kernel is the implicit-GEMM conv
            t_mid = torch.full((B,), 500.0, device=dev)
            model(noise_d, t_mid, context=cond_d)
            with ops.profile() as rec:
                model(noise_d, t_mid, context=cond_d)
            by = {}
            for tag, work, ms in rec.rows:
                a = by.setdefault(tag, [0.0, 0.0, 0])
                a[0] += work; a[1] += ms; a[2] += 1
            # dominant kernel = the conv variant with the largest share of the forward (the rolling-row kernel with
            # the fused GroupNorm operand transform on LDCT-512)
            conv_tags = [k for k in by if k.startswith("conv_rolling") or k == "conv_tile"]
            top = max(conv_tags, key=lambda k: by[k][1])
            cw, cms, cn = by[top]
            achieved = cw / (cms * 1e-3) / 1e12
            fwd_ms = sum(v[1] for v in by.values())
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
                "flop_per_launch": cw / cn, "avg_launch_ms": cms / cn, "launches_per_forward": cn,
                "share_of_forward": cms / fwd_ms,
                "frac_of_burst_peak": achieved / peaks["tf_burst"],
                "all_conv_kernels": {"achieved": allw / (allms * 1e-3) / 1e12, "share_of_forward": allms / fwd_ms,
                                     "launches_per_forward": sum(by[k][2] for k in conv_tags)},
                "groupnorm": {"bound": "hbm", "achieved": gn[0] / (gn[1] * 1e-3) / 1e9, "peak": peaks["hbm"],
                              "unit": "GB/s", "frac": gn[0] / (gn[1] * 1e-3) / 1e9 / peaks["hbm"],
                              "share_of_forward": gn[1] / fwd_ms,
                              "note": "standalone GroupNorm apply (small levels, attention norms) only: on rows >= 65 "
                                      "px the apply runs inside the consumer conv; algorithmic bytes = 1 read + 1 "
                                      "write bf16"},
                "per_kernel_ms_per_forward": {k: round(v[1], 3) for k, v in by.items()},
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

    if rank == 0:
        value = total * args.steps / elapsed
        e2e_value = total * args.steps / e2e_elapsed
        model_tflops = value * FLOP_PER_SAMPLE_FWD * EULER_STEPS / 1e12 / world
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
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
    ap.add_argument("--eager-baseline", action="store_true",
                    help="also time the reference's plain-PyTorch path eagerly on the GPU (context figure)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
