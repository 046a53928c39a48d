"""`run_model.py --mode {sample,evaluate}` for the B200 sampling path.

Mirror of the sampling / evaluation branches of the reference dispatcher (`src/run_model.py:31-106` ->
`DiffusionLikeSampler.sample` / `.evaluate`, `src/pipelines/samplers/diffusion_like.py:77-338`): same flags for the
parts that exist here (`--ckpt_dir --mode --batch_size --device --seed --num_samples --num_inference_steps
--start_step --last_n_steps --scheduler --save --output_dir`).  `evaluate` compares the clamped samples with targets
(`--targets_pt`), initialises from the noised target when `--start_step/--last_n_steps` are given (reference :229), and
writes `eval_metrics.csv` / `eval_metrics_per_image.csv` with the reference's columns (MSE, PSNR = 10 log10(1/MSE),
model_samples_per_second = count / sum of model time); SSIM needs scikit-image, which - as in the reference when it
is missing - is reported as unavailable.  The reference reads its conditioning images through its dataset classes (`--data_txt`), which are out
of scope; here conditioning comes from a tensor file (`--conditioning_pt`, a `(N, C, H, W)` tensor in [0, 1]) or is
synthetic LDCT-shaped noise (`--synthetic N H W`).  Multi-GPU: launch with torchrun; the sample index range is sharded
across ranks and gathered on rank 0.

  python -m fmdm_b200.run_model --ckpt_dir runs/ldct_fm --scheduler flowmatch --num_inference_steps 50 \
         --synthetic 16 512 512 --save
"""
from __future__ import annotations

import argparse
import json
import logging
import time
from pathlib import Path

import torch


def load_run_config(ckpt_dir: Path) -> dict:
    """`train_config.json` of a run directory (reference: `utils/sampling_utils.py::load_run_config`)."""
    path = Path(ckpt_dir) / "train_config.json"
    if not path.exists():
        raise FileNotFoundError(f"{path} not found (a checkpoint directory holds train_config.json)")
    with open(path) as f:
        return json.load(f)


def resolve_checkpoint(ckpt_dir: Path, model_type: str):
    """Best, then last checkpoint of the run (`{flow,diff}_{best,last}.pt`, `flow_matching_lib.py:197-211`)."""
    prefix = "flow" if str(model_type).lower() == "flow_matching" else "diff"
    for name in (f"{prefix}_best.pt", f"{prefix}_last.pt", f"{prefix}_best.safetensors", f"{prefix}_last.safetensors"):
        p = Path(ckpt_dir) / name
        if p.exists():
            return p
    return None


def evaluation_rows(generated: torch.Tensor, targets: torch.Tensor):
    """Per-image MSE / PSNR exactly as `DiffusionLikeSampler.evaluate` computes them (diffusion_like.py:246-249)."""
    generated, targets = generated.clamp(0.0, 1.0), targets.clamp(0.0, 1.0)
    dims = tuple(range(1, generated.ndim))
    mse = torch.mean((generated - targets) ** 2, dim=dims)
    psnr = 10.0 * torch.log10(1.0 / mse.clamp(min=1e-12))
    return mse, psnr


def write_eval_metrics(out_dir: Path, generated, targets, timing: dict, world: int = 1):
    mse, psnr = evaluation_rows(generated, targets)
    count = int(generated.shape[0])
    # timing holds this rank's model time for its shard; ranks run concurrently, so the job's model time is one shard's
    seconds = float(timing.get("model_seconds", 0.0))
    sps = count / seconds if seconds > 0 else 0.0
    row = {"samples": count, "mse": f"{float(mse.mean()):.8f}", "psnr": f"{float(psnr.mean()):.6f}", "ssim": "",
           "ssim_enabled": False, "model_seconds": f"{seconds:.6f}", "model_samples_per_second": f"{sps:.6f}",
           "model_seconds_per_sample": f"{(seconds / count if count else 0.0):.8f}",
           "model_calls": int(timing.get("model_calls", 0)) * world}
    with open(out_dir / "eval_metrics.csv", "w") as f:
        f.write(",".join(row.keys()) + "\n" + ",".join(str(v) for v in row.values()) + "\n")
    with open(out_dir / "eval_metrics_per_image.csv", "w") as f:
        f.write("sample_index,mse,psnr,ssim\n")
        for i in range(count):
            f.write(f"{i},{float(mse[i]):.8f},{float(psnr[i]):.6f},\n")
    print(f"Eval MSE: {float(mse.mean()):.6f} | PSNR: {float(psnr.mean()):.3f}")
    print(f"Model throughput: {sps:.3f} samples/s | model time {seconds:.3f}s")
    print("Eval SSIM: unavailable (install scikit-image)")
    return row


def main(argv=None) -> int:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(message)s", force=True)
    ap = argparse.ArgumentParser(description="Sample from a flow-matching / diffusion checkpoint on B200.")
    ap.add_argument("--ckpt_dir", type=Path, required=True)
    ap.add_argument("--mode", type=str, choices=("sample", "evaluate"), default="sample")
    ap.add_argument("--targets_pt", type=str, default=None, help="(N, C, H, W) target tensor file (evaluate)")
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--output_dir", type=str, default=None)
    ap.add_argument("--batch_size", type=int, default=4)
    ap.add_argument("--device", type=str, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--num_samples", type=int, default=None)
    ap.add_argument("--num_inference_steps", type=int, default=None)
    ap.add_argument("--start_step", type=int, default=None)
    ap.add_argument("--last_n_steps", type=int, default=None)
    ap.add_argument("--scheduler", type=str, default=None)
    ap.add_argument("--conditioning_pt", type=str, default=None, help="(N, C, H, W) conditioning tensor file")
    ap.add_argument("--synthetic", type=int, nargs=3, metavar=("N", "H", "W"), default=None)
    ap.add_argument("--random_init", action="store_true", help="run without a checkpoint (random-init weights)")
    args = ap.parse_args(argv)

    from .parallel import gather_samples, init_distributed, shard_bounds
    from .utils.model_utils import build_diffusion_model, decode_diffusion_batch

    rank, world, local_rank = init_distributed()
    device = torch.device(args.device) if args.device else torch.device("cuda", local_rank)
    cfg = load_run_config(args.ckpt_dir)
    model_type = cfg.get("model", {}).get("model_type", "flow_matching")
    if model_type not in ("flow_matching", "diffusion"):
        raise ValueError(f"model_type '{model_type}' is outside the sampling hot path")
    ckpt = None if args.random_init else resolve_checkpoint(args.ckpt_dir, model_type)
    if ckpt is None and not args.random_init:
        raise FileNotFoundError(f"no checkpoint in {args.ckpt_dir} (pass --random_init to sample from random weights)")
    torch.manual_seed(args.seed)
    model = build_diffusion_model(cfg, device, ckpt_path=ckpt)
    training_cfg, model_cfg = cfg["training"], cfg["model"]
    channels = int(training_cfg.get("channels", model_cfg.get("unet", {}).get("out_channels", 1)))

    g = torch.Generator().manual_seed(args.seed)
    if args.conditioning_pt:
        cond_all = torch.load(args.conditioning_pt, map_location="cpu", weights_only=True).float()
    elif args.synthetic:
        n, h, w = args.synthetic
        cond_all = (torch.rand(n, 1, h, w, generator=g) + 0.05 * torch.randn(n, 1, h, w, generator=g)).clamp_(0, 1)
    else:
        raise SystemExit("pass --conditioning_pt FILE or --synthetic N H W (the reference's dataset readers are out of scope)")
    targets_all = None
    if args.mode == "evaluate":
        if args.targets_pt:
            targets_all = torch.load(args.targets_pt, map_location="cpu", weights_only=True).float()
        elif args.synthetic:
            targets_all = cond_all.clone()  # synthetic smoke: "denoise towards the conditioning"
        else:
            raise SystemExit("--mode evaluate needs --targets_pt FILE")
        if targets_all.shape[0] != cond_all.shape[0]:
            raise SystemExit("targets and conditioning must hold the same number of images")
    if args.num_samples is not None:
        cond_all = cond_all[: args.num_samples]
        targets_all = None if targets_all is None else targets_all[: args.num_samples]
    total = cond_all.shape[0]
    noise_all = torch.randn(total, channels, *cond_all.shape[2:], generator=g)  # one global stream: results independent of N
    lo, hi = shard_bounds(total, rank, world)
    outs, timing = [], {}
    t0 = time.perf_counter()
    with torch.no_grad():
        for b0 in range(lo, hi, args.batch_size):
            b1 = min(b0 + args.batch_size, hi)
            cond = cond_all[b0:b1].to(device)
            # the reference draws its initial noise on the device inside the loop; here it comes from one global
            # CPU stream so the samples do not depend on the number of GPUs
            partial = (args.start_step is not None) or (args.last_n_steps is not None)
            ref = None if targets_all is None else targets_all[b0:b1].to(device)
            x = decode_diffusion_batch(model, training_cfg, model_cfg, device, tuple(noise_all[b0:b1].shape),
                                       conditioning_batch=cond, timing=timing,
                                       num_inference_steps=args.num_inference_steps, start_step=args.start_step,
                                       last_n_steps=args.last_n_steps, scheduler_override=args.scheduler,
                                       reference_batch=ref, init_from_reference=(ref is not None and partial),
                                       init_sample=noise_all[b0:b1].to(device))
            outs.append(x.clamp(0, 1))
    local = torch.cat(outs, 0) if outs else torch.empty((0, channels, *cond_all.shape[2:]), device=device)
    samples = gather_samples(local, total, rank, world)
    if rank == 0:
        wall = time.perf_counter() - t0
        calls = max(timing.get("model_calls", 0), 1)
        logging.info("sampled %d images in %.2f s (%.2f samples/s); model_samples_per_second %.2f", total, wall,
                     total / wall, (hi - lo) / max(timing.get("model_seconds", wall), 1e-9))
        out_dir = Path(args.output_dir or (args.ckpt_dir / "outputs")) / args.mode
        if args.save or args.mode == "evaluate":
            out_dir.mkdir(parents=True, exist_ok=True)
        if args.save:
            torch.save(samples.cpu(), out_dir / "samples.pt")
            logging.info("saved %s", out_dir / "samples.pt")
        if args.mode == "sample" and args.save:
            with open(out_dir / "eval_metrics.csv", "w") as f:
                f.write("count,model_calls,model_seconds,wall_seconds\n")
                f.write(f"{total},{calls},{timing.get('model_seconds', 0.0):.6f},{wall:.6f}\n")
        if args.mode == "evaluate":
            write_eval_metrics(out_dir, samples.cpu(), targets_all, timing, world)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
