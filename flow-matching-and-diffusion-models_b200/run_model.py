"""`run_model.py` for the B200 sampling path: the reference dispatcher's flag contract and handler seam
(`src/run_model.py:17-106`) over the modes that sit on the hot path.

* `main()` parses the reference's flags (`--ckpt_dir --mode --data_txt --save --output_dir --batch_size --device --seed
  --timestep --num_samples --num_inference_steps --start_step --last_n_steps --scheduler --save_input
  --save_conditioning --save_tensor_cache`), reads `model.model_type` from the run config, picks the handler class from
  `HANDLER_REGISTRY` and constructs it with the reference's sixteen keyword arguments (`run_model.py:75-92`); the
  reference's `tests/test_run_model_dispatch.py` passes against this module unchanged.
* `load_run_config` / `resolve_checkpoint`: `train_config.json`, else the legacy diffusers pipeline folder
  (`model_index.json` + `scheduler/scheduler_config.json` + `unet/config.json|txt`) and its
  `unet/diffusion_pytorch_model.safetensors` (`src/utils/sampling_utils.py:17-103, 106-167`).
* `DiffusionHandler` / `FlowMatchingHandler`: `sample`, `decode`, `evaluate`, `encode` of
  `src/pipelines/samplers/diffusion_like.py:31-338` (decode loop, add_noise initialisation for `--start_step /
  --last_n_steps`, per-image MSE / PSNR / SSIM, `eval_metrics.csv`, `eval_metrics_per_image.csv`, model-only
  throughput); `build_tensor_cache` / `debug_compare` and the `vae` handler belong to the dataset / autoencoder
  subsystems and raise `OutOfScopeError`.
* Data: the reference reads images through its dataset classes (split file `--data_txt`); those readers are out of
  scope, so images come from tensor files - `--data_txt bundle.pt` (a `torch.save`d dict with `image` = conditioning
  and / or `target`, each `(N, C, H, W)` in [0, 1]), or `--conditioning_pt` / `--targets_pt`, or `--synthetic N H W`
  (LDCT-shaped noise).  Any other `--data_txt` raises a clear error.
* SSIM is skimage's `structural_similarity` when scikit-image is installed, else the restatement below (same
  defaults: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance, border cropped), so `ssim_enabled` is True.
* Multi-GPU: launch with torchrun; the sample index range is sharded across ranks and gathered on rank 0.

  python -m fmdm_b200.run_model --ckpt_dir runs/ldct_fm --scheduler flowmatch --num_inference_steps 50 \
         --synthetic 16 512 512 --save
"""
from __future__ import annotations

import argparse
import json
import logging
import time
from pathlib import Path
from typing import Optional

import torch

from ._runtime import out_of_scope


# ----------------------------------------------------------------------------------------------------------------
# run config / checkpoint resolution (`src/utils/sampling_utils.py:17-167`)
# ----------------------------------------------------------------------------------------------------------------
_LEGACY_SCHEDULER_SKIP = ("_class_name", "_diffusers_version", "num_train_timesteps", "num_inference_steps",
                          "trained_betas")


def _legacy_diffusers_run_config(ckpt_dir: Path) -> dict:
    """Run config of a legacy diffusers pipeline folder (`sampling_utils.py:17-103`): the UNet hyper-parameters come
    from `unet/config.json` (or `.txt`), the scheduler from `scheduler/scheduler_config.json`; `in_channels` already
    counts the concatenated conditioning, hence `in_channels_already_conditioned`."""
    ckpt_dir = Path(ckpt_dir)
    index_path = ckpt_dir / "model_index.json"
    sched_path = ckpt_dir / "scheduler" / "scheduler_config.json"
    unet_path = ckpt_dir / "unet" / "config.json"
    if not unet_path.exists():
        unet_path = ckpt_dir / "unet" / "config.txt"
    if not (index_path.exists() and sched_path.exists() and unet_path.exists()):
        raise FileNotFoundError("Missing train_config.json and could not resolve a legacy diffusers folder layout.")
    index, sched, unet = (json.loads(p.read_text()) for p in (index_path, sched_path, unet_path))

    n_in, n_out = int(unet.get("in_channels", 1)), int(unet.get("out_channels", 1))
    conditioning = "concatenate" if n_in > n_out else None
    train_t = int(sched.get("num_train_timesteps", 1000))
    name = str(sched.get("_class_name", "DDPMScheduler")).replace("Scheduler", "").lower()
    unet_cfg = {"unet_impl": "diffusers_nd", "in_channels_already_conditioned": True,
                "sample_size": unet.get("sample_size", 256), "in_channels": n_in, "out_channels": n_out}
    for key, cast, default in (("layers_per_block", int, 2),
                               ("block_out_channels", tuple, [128, 128, 256, 256, 512, 512]),
                               ("down_block_types", tuple, []), ("up_block_types", tuple, []),
                               ("attention_head_dim", int, 8), ("norm_num_groups", int, 32), ("norm_eps", float, 1e-5),
                               ("flip_sin_to_cos", bool, True), ("freq_shift", int, 0),
                               ("center_input_sample", bool, False), ("resnet_time_scale_shift", str, "default"),
                               ("add_attention", bool, True)):
        unet_cfg[key] = cast(unet.get(key, default))
    return {
        "training": {"data_root": "/", "dataset": "ldct", "channels": n_out, "img_size": int(unet.get("sample_size", 256)),
                     "num_train_timesteps": train_t, "num_inference_steps": train_t, "conditioning": conditioning,
                     "load_ldct": conditioning in ("concatenate", "attention"), "norm": True},
        "model": {"model_type": "diffusion", "conditioning": conditioning,
                  "scheduler": {"name": name, "num_train_timesteps": train_t, "num_inference_steps": train_t,
                                "params": {k: v for k, v in sched.items() if k not in _LEGACY_SCHEDULER_SKIP}},
                  "unet": unet_cfg,
                  "legacy_source": {"model_index": index, "scheduler_config_path": str(sched_path),
                                    "unet_config_path": str(unet_path)}},
        "__config_path__": str(index_path),
    }


def load_run_config(ckpt_dir: Path) -> dict:
    """`train_config.json` of a run directory, else the legacy diffusers folder (`sampling_utils.py:106-130`)."""
    path = Path(ckpt_dir) / "train_config.json"
    if not path.exists():
        return _legacy_diffusers_run_config(Path(ckpt_dir))
    cfg = json.loads(path.read_text())
    known = cfg.get("__config_path__")
    if not (known and Path(known).exists()):
        cfg["__config_path__"] = str(path)
    return cfg


def resolve_checkpoint(ckpt_dir: Path, model_type: str) -> Path:
    """Best, then last checkpoint of the run (`sampling_utils.py:133-167`); a diffusion run also resolves the legacy
    `unet/diffusion_pytorch_model.safetensors`; unknown model types take the last `*.pt`.  Raises FileNotFoundError."""
    ckpt_dir = Path(ckpt_dir)
    kind = str(model_type).lower()
    names = {"vae": ("vae_best.pt", "vae_last.pt"), "diffusion": ("diff_best.pt", "diff_last.pt"),
             "flow_matching": ("flow_best.pt", "flow_last.pt")}.get(kind)
    for name in names or ():
        if (ckpt_dir / name).exists():
            return ckpt_dir / name
    if kind == "diffusion":
        legacy = ckpt_dir / "unet" / "diffusion_pytorch_model.safetensors"
        if legacy.exists():
            return legacy
    if names is None:
        found = sorted(ckpt_dir.glob("*.pt"))
        if found:
            return found[-1]
    raise FileNotFoundError(f"No checkpoint found in {ckpt_dir}")


# ----------------------------------------------------------------------------------------------------------------
# metrics (`diffusion_like.py:246-263`, `utils/evaluation_utils.py:64-91`)
# ----------------------------------------------------------------------------------------------------------------
def structural_similarity(a, b, *, data_range: float = 1.0, win_size: int = 7, k1: float = 0.01, k2: float = 0.03,
                          channel_axis=None) -> float:
    """Mean SSIM of two equally shaped N-d arrays (Wang et al. 2004) with scikit-image's defaults, which the reference
    calls as `ssim_fn(p, t, channel_axis=None, data_range=1.0)`: uniform `win_size` window, sample (N-1) covariance,
    float64, the half-window border cropped before averaging."""
    import numpy as np
    from scipy.ndimage import uniform_filter

    if channel_axis is not None:
        raise ValueError("structural_similarity: pass one channel at a time (channel_axis=None)")
    x, y = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if x.shape != y.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(x.shape) < win_size:
        raise ValueError(f"win_size {win_size} exceeds image extent {x.shape}")
    n = float(win_size ** x.ndim)
    unbiased = n / (n - 1.0)
    mx, my = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
    var_x = unbiased * (uniform_filter(x * x, size=win_size) - mx * mx)
    var_y = unbiased * (uniform_filter(y * y, size=win_size) - my * my)
    cov = unbiased * (uniform_filter(x * y, size=win_size) - mx * my)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2.0 * mx * my + c1) * (2.0 * cov + c2)) / ((mx * mx + my * my + c1) * (var_x + var_y + c2))
    pad = (win_size - 1) // 2
    return float(s[tuple(slice(pad, d - pad) for d in s.shape)].mean(dtype=np.float64))


def resolve_ssim_fn():
    try:
        from skimage.metrics import structural_similarity as fn  # the reference's import (`diffusion_like.py:169`)
        return fn
    except Exception:
        return structural_similarity


def compute_ssim_sample(pred: torch.Tensor, tgt: torch.Tensor, ssim_fn) -> Optional[float]:
    """SSIM of one channel-first sample, averaged over channels (`evaluation_utils.py:64-91`)."""
    if pred.shape != tgt.shape or pred.ndim < 2:
        return None
    p, t = pred.detach().cpu().float(), tgt.detach().cpu().float()
    if p.ndim == 2:
        return float(ssim_fn(p.numpy(), t.numpy(), channel_axis=None, data_range=1.0))
    scores = [float(ssim_fn(p[c].numpy(), t[c].numpy(), channel_axis=None, data_range=1.0))
              for c in range(p.shape[0]) if p[c].ndim >= 2]
    return sum(scores) / len(scores) if scores else None


def evaluation_rows(generated: torch.Tensor, targets: torch.Tensor):
    """Per-image MSE / PSNR exactly as `DiffusionLikeSampler.evaluate` computes them (diffusion_like.py:246-249)."""
    generated, targets = generated.clamp(0.0, 1.0), targets.clamp(0.0, 1.0)
    dims = tuple(range(1, generated.ndim))
    mse = torch.mean((generated - targets) ** 2, dim=dims)
    psnr = 10.0 * torch.log10(1.0 / mse.clamp(min=1e-12))
    return mse, psnr


def write_eval_metrics(out_dir: Path, generated, targets, timing: dict, world: int = 1, indices=None):
    """`eval_metrics.csv` (one row, the reference's columns) and `eval_metrics_per_image.csv`."""
    mse, psnr = evaluation_rows(generated, targets)
    count = int(generated.shape[0])
    ssim_fn = resolve_ssim_fn()
    g, t = generated.clamp(0.0, 1.0), targets.clamp(0.0, 1.0)
    ssim_values = []
    for i in range(count):
        try:
            ssim_values.append(compute_ssim_sample(g[i], t[i], ssim_fn))
        except ValueError:          # image smaller than the 7-pixel window
            ssim_values.append(None)
    have = [v for v in ssim_values if v is not None]
    avg_ssim = sum(have) / len(have) if have else None
    # timing holds this rank's model time for its shard; ranks run concurrently, so the job's model time is one shard's
    seconds = float(timing.get("model_seconds", 0.0))
    sps = count / seconds if seconds > 0 else 0.0
    row = {"samples": count, "mse": f"{float(mse.mean()):.8f}", "psnr": f"{float(psnr.mean()):.6f}",
           "ssim": "" if avg_ssim is None else f"{avg_ssim:.6f}", "ssim_enabled": True,
           "model_seconds": f"{seconds:.6f}", "model_samples_per_second": f"{sps:.6f}",
           "model_seconds_per_sample": f"{(seconds / count if count else 0.0):.8f}",
           "model_calls": int(timing.get("model_calls", 0)) * world}
    with open(out_dir / "eval_metrics.csv", "w") as f:
        f.write(",".join(row.keys()) + "\n" + ",".join(str(v) for v in row.values()) + "\n")
    with open(out_dir / "eval_metrics_per_image.csv", "w") as f:
        f.write("sample_index,mse,psnr,ssim\n")
        for i in range(count):
            s = "" if ssim_values[i] is None else f"{ssim_values[i]:.6f}"
            f.write(f"{i if indices is None else indices[i]},{float(mse[i]):.8f},{float(psnr[i]):.6f},{s}\n")
    print(f"Eval MSE: {float(mse.mean()):.6f} | PSNR: {float(psnr.mean()):.3f}")
    print(f"Model throughput: {sps:.3f} samples/s | model time {seconds:.3f}s")
    if avg_ssim is not None:
        print(f"Eval SSIM: {avg_ssim:.4f}")
    return row


# ----------------------------------------------------------------------------------------------------------------
# tensor-file data source (stands in for the reference's dataset readers)
# ----------------------------------------------------------------------------------------------------------------
class TensorSource:
    """Conditioning (`image`) and target tensors of a run, `(N, C, H, W)` float in [0, 1]."""

    def __init__(self, image: Optional[torch.Tensor], target: Optional[torch.Tensor]):
        self.image, self.target = image, target
        both = [t for t in (image, target) if t is not None]
        if not both:
            raise ValueError("TensorSource needs conditioning and / or target tensors")
        if len({t.shape[0] for t in both}) != 1:
            raise ValueError("targets and conditioning must hold the same number of images")

    def __len__(self) -> int:
        return int((self.image if self.image is not None else self.target).shape[0])

    def take(self, count: Optional[int]) -> "TensorSource":
        if count is None:
            return self
        return TensorSource(None if self.image is None else self.image[:count],
                            None if self.target is None else self.target[:count])

    @staticmethod
    def resolve(data_txt, conditioning_pt, targets_pt, synthetic, seed: int) -> "TensorSource":
        def load(path):
            return torch.load(path, map_location="cpu", weights_only=True)

        if data_txt is not None:
            if not str(data_txt).endswith((".pt", ".pth")):
                out_of_scope("dataset split files (--data_txt): pass a tensor bundle (.pt with 'image' / 'target'), "
                             "--conditioning_pt / --targets_pt, or --synthetic N H W")
            bundle = load(data_txt)
            if not isinstance(bundle, dict):
                raise ValueError("--data_txt tensor bundle must be a dict with 'image' and / or 'target'")
            img, tgt = bundle.get("image"), bundle.get("target")
            return TensorSource(None if img is None else img.float(), None if tgt is None else tgt.float())
        if conditioning_pt or targets_pt:
            return TensorSource(load(conditioning_pt).float() if conditioning_pt else None,
                                load(targets_pt).float() if targets_pt else None)
        if synthetic:
            n, h, w = synthetic
            g = torch.Generator().manual_seed(seed)
            img = (torch.rand(n, 1, h, w, generator=g) + 0.05 * torch.randn(n, 1, h, w, generator=g)).clamp_(0, 1)
            return TensorSource(img, img.clone())   # synthetic smoke: "denoise towards the conditioning"
        raise SystemExit("pass --data_txt BUNDLE.pt, --conditioning_pt FILE or --synthetic N H W "
                         "(the reference's dataset readers are out of scope)")


# ----------------------------------------------------------------------------------------------------------------
# handlers (`src/pipelines/samplers/handlers/{base,diffusion_handler,flow_matching_handler,vae_handler}.py`)
# ----------------------------------------------------------------------------------------------------------------
class ModelHandler:
    """Same constructor as the reference's `ModelHandler` (`handlers/base.py:17-66`); `extras` carries this repo's
    tensor-file data flags (`conditioning_pt`, `targets_pt`, `synthetic`, `random_init`)."""

    model_type = None

    def __init__(self, ckpt_dir, data_txt=None, save=False, output_dir=None, batch_size=4, device=None, seed=42,
                 timestep=None, num_samples=None, save_input=False, save_conditioning=False, num_inference_steps=None,
                 start_step=None, last_n_steps=None, scheduler=None, save_tensor_cache=False, **extras):
        self.ckpt_dir = Path(ckpt_dir)
        self.data_txt = data_txt
        self.save = save
        self.output_dir = output_dir
        self.batch_size = batch_size
        self.device = device
        self.seed = seed
        self.timestep = timestep
        self.num_samples = num_samples
        self.save_input = save_input
        self.save_conditioning = save_conditioning
        self.num_inference_steps = num_inference_steps
        self.start_step = start_step
        self.last_n_steps = last_n_steps
        self.scheduler = scheduler
        self.save_tensor_cache = bool(save_tensor_cache)
        self.extras = extras

    def encode(self) -> None:
        raise NotImplementedError

    def decode(self) -> None:
        raise NotImplementedError

    def sample(self) -> None:
        raise NotImplementedError

    def evaluate(self) -> None:
        raise NotImplementedError

    def build_tensor_cache(self) -> None:
        out_of_scope("run_model --mode build_tensor_cache (dataset tensor caches)")

    def debug_compare(self) -> None:
        out_of_scope("run_model --mode debug_compare")


class _DiffusionLikeHandler(ModelHandler):
    """`DiffusionLikeSampler` (`src/pipelines/samplers/diffusion_like.py`) over tensor-file data."""

    # ---- shared set-up ----------------------------------------------------------------------------------------
    def _setup(self, need_model: bool = True):
        from .parallel import init_distributed
        from .utils.model_utils import build_diffusion_model

        self.rank, self.world, local_rank = init_distributed()
        self.dev = torch.device(self.device) if self.device else torch.device("cuda", local_rank)
        self.cfg = load_run_config(self.ckpt_dir)
        self.training_cfg, self.model_cfg = self.cfg["training"], self.cfg["model"]
        self.channels = int(self.training_cfg.get("channels", self.model_cfg.get("unet", {}).get("out_channels", 1)))
        torch.manual_seed(self.seed)
        self.model = None
        if need_model:
            ckpt = None if self.extras.get("random_init") else resolve_checkpoint(self.ckpt_dir, self.model_type)
            self.model = build_diffusion_model(self.cfg, self.dev, ckpt_path=ckpt)
        self.data = TensorSource.resolve(self.data_txt, self.extras.get("conditioning_pt"),
                                         self.extras.get("targets_pt"), self.extras.get("synthetic"),
                                         self.seed).take(self.num_samples)

    def _out_dir(self, mode: str) -> Path:
        return Path(self.output_dir or (self.ckpt_dir / "outputs")) / mode

    def _decode_all(self, mode: str, use_targets: bool):
        """The batch loop of `_run_decode` / `_run_evaluate` (`diffusion_like.py:114-139, 217-237`): returns the
        clamped samples of the whole index range (rank 0), the timing dict and the wall time."""
        from .parallel import gather_samples, shard_bounds
        from .pipelines.utils import resolve_conditioning_mode
        from .utils.model_utils import decode_diffusion_batch

        data, dev = self.data, self.dev
        total = len(data)
        mode_c = resolve_conditioning_mode(self.training_cfg.get("conditioning") or self.model_cfg.get("conditioning"))
        spatial = tuple((data.image if data.image is not None else data.target).shape[2:])
        g = torch.Generator().manual_seed(self.seed)
        # the reference draws its initial noise on the device inside the loop; here it comes from one global CPU
        # stream so the samples do not depend on the number of GPUs or on the batch size
        noise_all = torch.randn(total, self.channels, *spatial, generator=g)
        lo, hi = shard_bounds(total, self.rank, self.world)
        partial = (self.start_step is not None) or (self.last_n_steps is not None)
        outs, timing = [], {"model_seconds": 0.0, "model_calls": 0}
        t0 = time.perf_counter()
        with torch.no_grad():
            for b0 in range(lo, hi, self.batch_size):
                b1 = min(b0 + self.batch_size, hi)
                cond = None
                if mode_c in ("concatenate", "attention") and data.image is not None:
                    cond = data.image[b0:b1].to(dev)
                ref = data.target[b0:b1].to(dev) if (use_targets and data.target is not None) else None
                x = decode_diffusion_batch(self.model, self.training_cfg, self.model_cfg, dev,
                                           tuple(noise_all[b0:b1].shape), cond, timing=timing,
                                           num_inference_steps=self.num_inference_steps, start_step=self.start_step,
                                           last_n_steps=self.last_n_steps, scheduler_override=self.scheduler,
                                           reference_batch=ref, init_from_reference=(ref is not None and partial),
                                           init_sample=noise_all[b0:b1].to(dev))
                outs.append(x.clamp(0.0, 1.0))
        local = torch.cat(outs, 0) if outs else torch.empty((0, self.channels, *spatial), device=dev)
        samples = gather_samples(local, total, self.rank, self.world)
        wall = time.perf_counter() - t0
        if self.rank == 0:
            logging.info("%s %s: %d images in %.2f s (%.2f samples/s); model_samples_per_second %.2f",
                         self.model_type, mode, total, wall, total / max(wall, 1e-9),
                         (hi - lo) / max(timing.get("model_seconds", wall), 1e-9))
        return samples, timing, wall

    def _save_tensors(self, out_dir: Path, samples: torch.Tensor) -> None:
        out_dir.mkdir(parents=True, exist_ok=True)
        torch.save(samples.cpu(), out_dir / "samples.pt")
        logging.info("saved %s", out_dir / "samples.pt")
        if self.save_input and self.data.target is not None:           # `diffusion_like.py:141-142, 241-242`
            torch.save(self.data.target, out_dir / "input.pt")
        if self.save_conditioning and self.data.image is not None:     # `:143-144, 243-244`
            torch.save(self.data.image, out_dir / "conditioning.pt")

    # ---- modes ------------------------------------------------------------------------------------------------
    def decode(self) -> None:
        self._setup()
        samples, timing, wall = self._decode_all("decode", use_targets=True)
        if self.rank == 0 and self.save:
            self._save_tensors(self._out_dir("decode"), samples)

    def sample(self) -> None:
        self._setup()
        samples, timing, wall = self._decode_all("sample", use_targets=False)
        if self.rank == 0 and self.save:
            out_dir = self._out_dir("sample")
            self._save_tensors(out_dir, samples)
            with open(out_dir / "eval_metrics.csv", "w") as f:
                f.write("count,model_calls,model_seconds,wall_seconds\n")
                f.write(f"{len(self.data)},{max(timing.get('model_calls', 0), 1)},"
                        f"{timing.get('model_seconds', 0.0):.6f},{wall:.6f}\n")

    def evaluate(self) -> None:
        self._setup()
        if self.data.target is None:
            raise SystemExit("--mode evaluate needs targets (--targets_pt FILE or a bundle with 'target')")
        samples, timing, wall = self._decode_all("evaluate", use_targets=True)
        if self.rank == 0:
            out_dir = self._out_dir("evaluate")
            out_dir.mkdir(parents=True, exist_ok=True)
            if self.save:
                self._save_tensors(out_dir, samples)
            write_eval_metrics(out_dir, samples.cpu(), self.data.target, timing, self.world)
            run_cfg = {"mode": "evaluate", "model_type": self.model_type, "ckpt_dir": str(self.ckpt_dir),
                       "data_txt": self.data_txt, "scheduler": self.scheduler,
                       "num_inference_steps": self.num_inference_steps, "start_step": self.start_step,
                       "last_n_steps": self.last_n_steps, "num_samples": self.num_samples,
                       "batch_size": self.batch_size, "seed": self.seed, "save": self.save,
                       "save_input": self.save_input, "save_conditioning": self.save_conditioning}
            (out_dir / "run_config.json").write_text(json.dumps(run_cfg, indent=2))

    def encode(self) -> None:
        """`_run_encode` (`diffusion_like.py:31-74`): forward-noise the targets to `--timestep` (random per image
        when absent) with the configured scheduler's `add_noise`."""
        from .pipelines.utils import build_scheduler
        from .utils.model_utils import encode_diffusion_batch

        self._setup(need_model=False)
        if self.data.target is None:
            raise SystemExit("--mode encode needs targets (--targets_pt FILE or a bundle with 'target')")
        scheduler, _ = build_scheduler(self.model_cfg.get("scheduler", {}), self.training_cfg)
        if not hasattr(scheduler, "add_noise"):
            raise ValueError(f"scheduler '{type(scheduler).__name__}' has no add_noise; encode needs a diffusion scheduler")
        outs = []
        for b0 in range(0, len(self.data), self.batch_size):
            tgt = self.data.target[b0:b0 + self.batch_size].to(self.dev)
            if self.timestep is None:
                ts = torch.randint(0, scheduler.config.num_train_timesteps, (tgt.size(0),), device=self.dev).long()
            else:
                ts = torch.full((tgt.size(0),), int(self.timestep), device=self.dev, dtype=torch.long)
            outs.append(encode_diffusion_batch(scheduler, tgt, ts).cpu())
        if self.rank == 0 and self.save:
            out_dir = self._out_dir("encode")
            out_dir.mkdir(parents=True, exist_ok=True)
            torch.save(torch.cat(outs, 0), out_dir / "encoded.pt")
        logging.info("%s encode completed for %d samples.", self.model_type, len(self.data))


class DiffusionHandler(_DiffusionLikeHandler):
    model_type = "diffusion"


class FlowMatchingHandler(_DiffusionLikeHandler):
    model_type = "flow_matching"


class VAEHandler(ModelHandler):
    """Autoencoder workflows are outside the sampling hot path (only `AutoencoderKL.decode` is
    built, as a library call); every mode raises `OutOfScopeError`."""

    model_type = "vae"

    def _refuse(self):
        out_of_scope("run_model for model_type 'vae' (autoencoder encode / decode / evaluate workflows)")

    encode = decode = sample = evaluate = _refuse


HANDLER_REGISTRY = {
    "vae": VAEHandler,
    "diffusion": DiffusionHandler,
    "flow_matching": FlowMatchingHandler,
}


def _resolve_handler(model_type: str):
    key = str(model_type).lower()
    if key not in HANDLER_REGISTRY:
        raise ValueError(f"Unsupported model_type '{model_type}'.")
    return HANDLER_REGISTRY[key]


MODES = ("sample", "encode", "decode", "evaluate", "build_tensor_cache", "debug_compare")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Run sampling/encoding/decoding/eval from a checkpoint dir on B200.")
    ap.add_argument("--ckpt_dir", type=Path, required=True, help="Checkpoint directory containing train_config.json.")
    ap.add_argument("--mode", type=str, choices=MODES, default="sample")
    ap.add_argument("--data_txt", type=str, default=None,
                    help="Tensor bundle (.pt dict with 'image' / 'target'); dataset split files are out of scope.")
    ap.add_argument("--save", action="store_true", help="Save outputs to disk.")
    ap.add_argument("--output_dir", type=str, default=None, help="Output root (defaults to ckpt_dir/outputs).")
    ap.add_argument("--batch_size", type=int, default=4)
    ap.add_argument("--device", type=str, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--timestep", type=int, default=None, help="Optional timestep for encode.")
    ap.add_argument("--num_samples", type=int, default=None)
    ap.add_argument("--num_inference_steps", type=int, default=None)
    ap.add_argument("--start_step", type=int, default=None,
                    help="Start denoising from this train-timestep index (e.g. 700 runs from t<=700).")
    ap.add_argument("--last_n_steps", type=int, default=None, help="Run only the last N denoising steps.")
    ap.add_argument("--scheduler", type=str, default=None,
                    help="ddpm, ddim, dpmsolver1, dpmsolver2, dpmsolver++, dpmsolversde, unipc, flowmatch")
    ap.add_argument("--save_input", action="store_true", help="Also save model inputs when --save is enabled.")
    ap.add_argument("--save_conditioning", action="store_true", help="Also save conditioning tensors with --save.")
    ap.add_argument("--save_tensor_cache", action="store_true", help="(dataset tensor caches; accepted, unused here)")
    # tensor-file data flags of this repo (the reference's dataset readers are out of scope)
    ap.add_argument("--conditioning_pt", type=str, default=None, help="(N, C, H, W) conditioning tensor file")
    ap.add_argument("--targets_pt", type=str, default=None, help="(N, C, H, W) target tensor file")
    ap.add_argument("--synthetic", type=int, nargs=3, metavar=("N", "H", "W"), default=None)
    ap.add_argument("--random_init", action="store_true", help="run without a checkpoint (random-init weights)")
    return ap


def main(argv=None) -> int:
    """Dispatch a model workflow from a checkpoint directory (`src/run_model.py:31-106`)."""
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(message)s", force=True)
    args = build_parser().parse_args(argv)
    cfg = load_run_config(args.ckpt_dir)
    handler_cls = _resolve_handler(cfg.get("model", {}).get("model_type", "vae"))
    extras = {k: v for k, v in (("conditioning_pt", args.conditioning_pt), ("targets_pt", args.targets_pt),
                                ("synthetic", args.synthetic), ("random_init", args.random_init)) if v}
    handler = handler_cls(
        ckpt_dir=args.ckpt_dir, data_txt=args.data_txt, save=args.save, output_dir=args.output_dir,
        batch_size=args.batch_size, device=args.device, seed=args.seed, timestep=args.timestep,
        num_samples=args.num_samples, save_input=args.save_input, save_conditioning=args.save_conditioning,
        num_inference_steps=args.num_inference_steps, start_step=args.start_step, last_n_steps=args.last_n_steps,
        scheduler=args.scheduler, save_tensor_cache=args.save_tensor_cache, **extras)
    with torch.no_grad():
        getattr(handler, args.mode if args.mode in MODES else "sample")()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
