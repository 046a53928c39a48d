"""Host logic directly above the hot loop (SURVEY.md §8a row a17 and §8b checkpoint contract): checkpoint loading with
the legacy diffusers key remap, `decode_diffusion_batch`'s scheduler override / timestep subset / add_noise init.
CPU part: everything that needs no kernel.  GPU part: decode against the oracle loop."""
import json

import pytest
import torch

from fmdm_b200.models.generators import DiffusionUNetFactory
from fmdm_b200.utils.model_utils import diffusion_utils as DU

SMALL = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
         "block_out_channels": [32, 64], "down_block_types": ["DownBlock2D", "AttnDownBlock2D"],
         "up_block_types": ["AttnUpBlock2D", "UpBlock2D"]}
CFG = {"model": {"model_type": "flow_matching", "unet": SMALL, "conditioning": "concatenate",
                 "scheduler": {"name": "flow_match_euler", "params": {}}},
       "training": {"channels": 1, "conditioning": "concatenate", "num_train_timesteps": 1000}}

# inverse of the legacy rules: how a diffusers-era checkpoint names the same tensors
_TO_LEGACY = [(".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn."),
              (".conv1.conv.", ".conv1."), (".conv2.conv.", ".conv2."), (".emb_layers.", ".time_emb_proj."),
              (".skip_connection.conv.", ".conv_shortcut."), (".downsamplers.0.op.conv.", ".downsamplers.0.conv."),
              (".upsamplers.0.conv.conv.", ".upsamplers.0.conv.")]


def _legacy_names(sd):
    out = {}
    for k, v in sd.items():
        for a, b in _TO_LEGACY:
            k = k.replace(a, b)
        out[k] = v
    return out


def test_checkpoint_roundtrip_and_legacy_remap(tmp_path):
    torch.manual_seed(3)
    ref = DiffusionUNetFactory().build(SMALL, "concatenate", 1)
    sd = {k: torch.randn_like(v) for k, v in ref.state_dict().items()}
    # native checkpoint layout written by the reference trainers: {"model": state_dict, ...}
    p = tmp_path / "flow_best.pt"
    torch.save({"model": sd, "epoch": 3}, p)
    m = DU.build_diffusion_model(CFG, torch.device("cpu"), ckpt_path=p)
    assert not m.training
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # diffusers-era names: falls back to the remap automatically, and explicitly with load_legacy
    legacy = _legacy_names(sd)
    assert set(legacy) != set(sd)
    p2 = tmp_path / "legacy.pt"
    torch.save(legacy, p2)
    for load_legacy in (False, True):
        cfg = json.loads(json.dumps(CFG))
        cfg["model"]["unet"]["load_legacy"] = load_legacy
        m2 = DU.build_diffusion_model(cfg, torch.device("cpu"), ckpt_path=p2)
        for k, v in m2.state_dict().items():
            assert torch.equal(v, sd[k]), k
    # a shape mismatch is an error in strict mode, tolerated otherwise
    bad = dict(legacy)
    some = next(k for k in bad if k.endswith("conv_in.weight"))
    bad[some] = torch.zeros(7, 7)
    torch.save(bad, p2)
    with pytest.raises(RuntimeError, match="shape mismatches"):
        DU.build_diffusion_model(CFG, torch.device("cpu"), ckpt_path=p2)
    cfg = json.loads(json.dumps(CFG))
    cfg["model"]["unet"].update(load_legacy=True, legacy_strict_shapes=False)
    DU.build_diffusion_model(cfg, torch.device("cpu"), ckpt_path=p2)
    # missing tensors are reported in strict mode
    short = {k: v for k, v in legacy.items() if "mid_block" not in k}
    torch.save(short, p2)
    with pytest.raises(RuntimeError, match="key mismatch"):
        DU.build_diffusion_model(CFG, torch.device("cpu"), ckpt_path=p2)


def test_encode_and_conditioning_warning():
    class Sched:
        def add_noise(self, x, n, t):
            return x + 0.0 * n

    x = torch.ones(2, 1, 4, 4)
    assert torch.equal(DU.encode_diffusion_batch(Sched(), x, torch.zeros(2, dtype=torch.long)), x)
    assert DU.warn_attention_conditioning_shape(torch.zeros(2, 4, 8, 8), {"unet": {"cross_attention_dim": 8}}) is True
    assert DU.warn_attention_conditioning_shape(torch.zeros(2, 8, 8, 8), {"unet": {"cross_attention_dim": 8}}) is False
    assert DU.warn_attention_conditioning_shape(None, {"unet": {"cross_attention_dim": 8}}) is False
    assert DU.warn_attention_conditioning_shape(torch.zeros(2, 4, 8, 8), {"unet": {}}) is False


def test_run_model_config_helpers(tmp_path):
    from fmdm_b200 import run_model as RM

    with pytest.raises(FileNotFoundError):
        RM.load_run_config(tmp_path)
    (tmp_path / "train_config.json").write_text(json.dumps(CFG))
    cfg = RM.load_run_config(tmp_path)
    assert cfg["model"]["model_type"] == "flow_matching" and cfg["__config_path__"].endswith("train_config.json")
    with pytest.raises(FileNotFoundError):
        RM.resolve_checkpoint(tmp_path, "flow_matching")
    (tmp_path / "flow_last.pt").write_bytes(b"x")
    assert RM.resolve_checkpoint(tmp_path, "flow_matching").name == "flow_last.pt"
    (tmp_path / "flow_best.pt").write_bytes(b"x")
    assert RM.resolve_checkpoint(tmp_path, "flow_matching").name == "flow_best.pt"
    with pytest.raises(FileNotFoundError):
        RM.resolve_checkpoint(tmp_path, "diffusion")
    assert RM.resolve_checkpoint(tmp_path, "something_else").name == "flow_last.pt"   # last *.pt of the directory
    (tmp_path / "unet").mkdir()
    (tmp_path / "unet" / "diffusion_pytorch_model.safetensors").write_bytes(b"x")
    assert RM.resolve_checkpoint(tmp_path, "diffusion").name == "diffusion_pytorch_model.safetensors"


def test_legacy_diffusers_run_config(tmp_path):
    """A diffusers pipeline folder without train_config.json (`sampling_utils.py:17-103`): the run config is rebuilt
    from model_index / scheduler_config / unet config, builds the same module tree, and the safetensors checkpoint in
    diffusers' own key names loads through the legacy remap."""
    from safetensors.torch import save_file

    from fmdm_b200 import run_model as RM

    (tmp_path / "scheduler").mkdir()
    (tmp_path / "unet").mkdir()
    (tmp_path / "model_index.json").write_text(json.dumps({"_class_name": "DDPMPipeline"}))
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(
        {"_class_name": "DDPMScheduler", "_diffusers_version": "0.14.0", "num_train_timesteps": 1000,
         "beta_start": 1e-4, "beta_end": 0.02, "beta_schedule": "linear", "trained_betas": None}))
    unet = {"sample_size": 32, "in_channels": 2, "out_channels": 1, "layers_per_block": 1, "block_out_channels": [32, 64],
            "down_block_types": ["DownBlock2D", "AttnDownBlock2D"], "up_block_types": ["AttnUpBlock2D", "UpBlock2D"]}
    (tmp_path / "unet" / "config.txt").write_text(json.dumps(unet))   # the .txt spelling is accepted too
    cfg = RM.load_run_config(tmp_path)
    assert cfg["model"]["model_type"] == "diffusion" and cfg["model"]["conditioning"] == "concatenate"
    assert cfg["training"]["load_ldct"] is True and cfg["training"]["channels"] == 1 and cfg["training"]["img_size"] == 32
    sched = cfg["model"]["scheduler"]
    assert sched["name"] == "ddpm" and sched["params"] == {"beta_start": 1e-4, "beta_end": 0.02, "beta_schedule": "linear"}
    u = cfg["model"]["unet"]
    assert u["in_channels_already_conditioned"] is True and u["in_channels"] == 2 and u["attention_head_dim"] == 8
    assert u["flip_sin_to_cos"] is True and u["block_out_channels"] == (32, 64)
    # same module tree as the native config with in_channels 1 + concatenated conditioning
    torch.manual_seed(4)
    native = DiffusionUNetFactory().build(SMALL, "concatenate", 1)
    save_file(_legacy_names({k: v.contiguous() for k, v in native.state_dict().items()}),
              str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    ckpt = RM.resolve_checkpoint(tmp_path, cfg["model"]["model_type"])
    model = DU.build_diffusion_model(cfg, torch.device("cpu"), ckpt_path=ckpt)
    got = model.state_dict()
    assert set(got) == set(native.state_dict())
    assert all(torch.equal(got[k], v) for k, v in native.state_dict().items())
    with pytest.raises(FileNotFoundError):
        RM.load_run_config(tmp_path / "scheduler")


class _DummyHandler:
    init_kwargs = None
    called = None

    def __init__(self, **kwargs):
        type(self).init_kwargs = kwargs

    def encode(self):
        type(self).called = "encode"

    def decode(self):
        type(self).called = "decode"

    def evaluate(self):
        type(self).called = "evaluate"

    def sample(self):
        type(self).called = "sample"


def test_run_model_dispatch(monkeypatch, tmp_path):
    """The reference's `tests/test_run_model_dispatch.py:28-66` against the mirror: flags are forwarded to the handler
    the registry names for `model.model_type`, under the reference's keyword names."""
    from fmdm_b200 import run_model as RM

    ckpt_dir = tmp_path / "run"
    ckpt_dir.mkdir()

    def fake_load(path):
        assert path == ckpt_dir
        return {"model": {"model_type": "diffusion"}, "training": {}}

    monkeypatch.setattr(RM, "load_run_config", fake_load)
    monkeypatch.setattr(RM, "HANDLER_REGISTRY", {"diffusion": _DummyHandler})
    monkeypatch.setattr("sys.argv", ["run_model.py", "--ckpt_dir", str(ckpt_dir), "--mode", "evaluate", "--batch_size",
                                     "64", "--num_samples", "128", "--save", "--save_input", "--save_conditioning"])
    RM.main()
    kw = _DummyHandler.init_kwargs
    assert _DummyHandler.called == "evaluate"
    assert kw["batch_size"] == 64 and kw["num_samples"] == 128
    assert kw["save"] is True and kw["save_input"] is True and kw["save_conditioning"] is True
    # exactly the reference's sixteen keyword arguments (`src/run_model.py:75-92`) when no tensor-file flag is given
    assert set(kw) == {"ckpt_dir", "data_txt", "save", "output_dir", "batch_size", "device", "seed", "timestep",
                       "num_samples", "save_input", "save_conditioning", "num_inference_steps", "start_step",
                       "last_n_steps", "scheduler", "save_tensor_cache"}
    RM.main(["--ckpt_dir", str(ckpt_dir), "--mode", "encode", "--timestep", "250", "--scheduler", "unipc",
             "--data_txt", "split.txt", "--save_tensor_cache"])
    kw = _DummyHandler.init_kwargs
    assert _DummyHandler.called == "encode" and kw["timestep"] == 250 and kw["scheduler"] == "unipc"
    assert kw["data_txt"] == "split.txt" and kw["save_tensor_cache"] is True
    monkeypatch.setattr(RM, "load_run_config", lambda p: {"model": {"model_type": "gan"}})
    with pytest.raises(ValueError, match="Unsupported model_type"):
        RM.main(["--ckpt_dir", str(ckpt_dir)])
    assert set(RM.HANDLER_REGISTRY) == {"diffusion"}      # monkeypatched view
    monkeypatch.undo()
    assert set(RM.HANDLER_REGISTRY) == {"vae", "diffusion", "flow_matching"}


def test_handlers_refuse_what_is_out_of_scope(tmp_path):
    from fmdm_b200 import run_model as RM
    from fmdm_b200._runtime import OutOfScopeError

    h = RM.FlowMatchingHandler(ckpt_dir=tmp_path)
    for mode in ("build_tensor_cache", "debug_compare"):
        with pytest.raises(OutOfScopeError):
            getattr(h, mode)()
    with pytest.raises(OutOfScopeError):
        RM.VAEHandler(ckpt_dir=tmp_path).sample()
    with pytest.raises(OutOfScopeError):      # a dataset split file: the readers are not part of the path
        RM.TensorSource.resolve("train_split.txt", None, None, None, 42)
    g = torch.Generator().manual_seed(0)
    torch.save({"image": torch.rand(3, 1, 8, 8, generator=g), "target": torch.rand(3, 1, 8, 8, generator=g)},
               tmp_path / "bundle.pt")
    src = RM.TensorSource.resolve(str(tmp_path / "bundle.pt"), None, None, None, 42)
    assert len(src) == 3 and len(src.take(2)) == 2 and src.take(None) is src


def _ssim_by_windows(x, y, win=7, k1=0.01, k2=0.03, r=1.0):
    """Brute force over every fully inside window: means, unbiased variances and covariance from numpy."""
    import numpy as np

    vals = []
    for i in range(x.shape[0] - win + 1):
        for j in range(x.shape[1] - win + 1):
            a, b = x[i:i + win, j:j + win].ravel(), y[i:i + win, j:j + win].ravel()
            c = np.cov(a, b, ddof=1)
            c1, c2 = (k1 * r) ** 2, (k2 * r) ** 2
            vals.append((2 * a.mean() * b.mean() + c1) * (2 * c[0, 1] + c2)
                        / ((a.mean() ** 2 + b.mean() ** 2 + c1) * (c[0, 0] + c[1, 1] + c2)))
    return float(np.mean(vals))


def test_ssim_restatement():
    """SSIM (`evaluation_utils.py:64-91` calls skimage's `structural_similarity(p, t, channel_axis=None,
    data_range=1.0)`; scikit-image is not installed here): the restatement against a brute-force window loop and the
    closed forms for identical and constant images."""
    import numpy as np

    from fmdm_b200 import run_model as RM

    rng = np.random.default_rng(0)
    x = rng.random((19, 23))
    y = np.clip(x + 0.1 * rng.standard_normal(x.shape), 0, 1)
    assert abs(RM.structural_similarity(x, x) - 1.0) < 1e-12
    assert abs(RM.structural_similarity(x, y) - _ssim_by_windows(x, y)) < 1e-10
    a, b = np.full((9, 9), 0.25), np.full((9, 9), 0.75)
    assert abs(RM.structural_similarity(a, b) - (2 * 0.25 * 0.75 + 1e-4) / (0.25 ** 2 + 0.75 ** 2 + 1e-4)) < 1e-12
    with pytest.raises(ValueError):
        RM.structural_similarity(x[:5], y[:5])
    # channel-first samples: mean over channels; mismatched shapes -> None (`evaluation_utils.py:69-70`)
    p = torch.tensor(np.stack([x, y]), dtype=torch.float32)
    t = torch.tensor(np.stack([y, y]), dtype=torch.float32)
    want = 0.5 * (RM.structural_similarity(p[0].numpy(), t[0].numpy()) + 1.0)
    assert abs(RM.compute_ssim_sample(p, t, RM.structural_similarity) - want) < 1e-6
    assert RM.compute_ssim_sample(p, t[:1], RM.structural_similarity) is None


@pytest.mark.gpu
def test_run_model_sample_cli(tmp_path):
    """`run_model --mode sample` end to end from a run directory: config + checkpoint -> samples.pt + eval_metrics.csv."""
    from fmdm_b200 import run_model as RM

    cfg = json.loads(json.dumps(CFG))
    cfg["model"]["unet"]["block_out_channels"] = [64, 128]
    (tmp_path / "train_config.json").write_text(json.dumps(cfg))
    torch.manual_seed(1)
    model = DiffusionUNetFactory().build(cfg["model"]["unet"], "concatenate", 1)
    torch.save({"model": model.state_dict()}, tmp_path / "flow_best.pt")
    rc = RM.main(["--ckpt_dir", str(tmp_path), "--synthetic", "5", "32", "32", "--batch_size", "2",
                  "--num_inference_steps", "6", "--scheduler", "flowmatch", "--save"])
    assert rc == 0
    out = torch.load(tmp_path / "outputs" / "sample" / "samples.pt", weights_only=True)
    assert out.shape == (5, 1, 32, 32) and torch.isfinite(out).all() and float(out.min()) >= 0 and float(out.max()) <= 1
    assert (tmp_path / "outputs" / "sample" / "eval_metrics.csv").read_text().startswith("count,model_calls")
    # same seed -> same samples, independent of the batch size
    rc = RM.main(["--ckpt_dir", str(tmp_path), "--synthetic", "5", "32", "32", "--batch_size", "5",
                  "--num_inference_steps", "6", "--scheduler", "flowmatch", "--save", "--output_dir",
                  str(tmp_path / "o2")])
    out2 = torch.load(tmp_path / "o2" / "sample" / "samples.pt", weights_only=True)
    assert torch.allclose(out, out2, atol=2e-2)


def test_evaluation_metrics_formulas(tmp_path):
    from fmdm_b200 import run_model as RM

    gen = torch.tensor([[[[0.5, 1.5], [0.0, -1.0]]], [[[0.25, 0.25], [0.25, 0.25]]]])
    tgt = torch.tensor([[[[0.5, 1.0], [0.5, 0.0]]], [[[0.25, 0.25], [0.25, 0.25]]]])
    mse, psnr = RM.evaluation_rows(gen, tgt)
    assert abs(float(mse[0]) - 0.0625) < 1e-7 and float(mse[1]) == 0.0            # clamped to [0, 1] first
    assert abs(float(psnr[0]) - 10 * torch.log10(torch.tensor(16.0)).item()) < 1e-4
    assert abs(float(psnr[1]) - 120.0) < 1e-3                                       # mse floor 1e-12
    row = RM.write_eval_metrics(tmp_path, gen, tgt, {"model_seconds": 2.0, "model_calls": 10})
    assert row["model_samples_per_second"] == "1.000000" and row["samples"] == 2
    head = (tmp_path / "eval_metrics.csv").read_text().splitlines()[0]
    assert head == ("samples,mse,psnr,ssim,ssim_enabled,model_seconds,model_samples_per_second,"
                    "model_seconds_per_sample,model_calls")
    assert len((tmp_path / "eval_metrics_per_image.csv").read_text().splitlines()) == 3


@pytest.mark.gpu
def test_run_model_evaluate_cli(tmp_path):
    from fmdm_b200 import run_model as RM

    cfg = json.loads(json.dumps(CFG))
    cfg["model"]["unet"]["block_out_channels"] = [64, 128]
    cfg["model"]["scheduler"] = {"name": "ddim", "params": {"beta_start": 1e-4, "beta_end": 0.02}}
    cfg["model"]["model_type"] = "diffusion"
    (tmp_path / "train_config.json").write_text(json.dumps(cfg))
    torch.manual_seed(2)
    model = DiffusionUNetFactory().build(cfg["model"]["unet"], "concatenate", 1)
    torch.save({"model": model.state_dict()}, tmp_path / "diff_last.pt")
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(4, 1, 32, 32, generator=g)
    tgt = torch.rand(4, 1, 32, 32, generator=g)
    torch.save(cond, tmp_path / "cond.pt")
    torch.save(tgt, tmp_path / "tgt.pt")
    rc = RM.main(["--ckpt_dir", str(tmp_path), "--mode", "evaluate", "--conditioning_pt", str(tmp_path / "cond.pt"),
                  "--targets_pt", str(tmp_path / "tgt.pt"), "--batch_size", "4", "--num_inference_steps", "10",
                  "--start_step", "300"])   # partial trajectory from the noised targets (add_noise init)
    assert rc == 0
    rows = (tmp_path / "outputs" / "evaluate" / "eval_metrics.csv").read_text().splitlines()
    vals = dict(zip(rows[0].split(","), rows[1].split(",")))
    assert vals["samples"] == "4" and float(vals["psnr"]) > 0 and float(vals["model_samples_per_second"]) > 0
    assert int(vals["model_calls"]) == 4  # ddim, 10 steps, leading spacing: timesteps <= 300 are 300, 200, 100, 0
    assert vals["ssim_enabled"] == "True" and -1.0 <= float(vals["ssim"]) <= 1.0
    per_image = (tmp_path / "outputs" / "evaluate" / "eval_metrics_per_image.csv").read_text().splitlines()
    assert len(per_image) == 5 and all(row.split(",")[3] != "" for row in per_image[1:])
    assert json.loads((tmp_path / "outputs" / "evaluate" / "run_config.json").read_text())["start_step"] == 300
    # the reference's flag names: a tensor bundle through --data_txt, inputs / conditioning saved next to the samples
    torch.save({"image": cond, "target": tgt}, tmp_path / "bundle.pt")
    rc = RM.main(["--ckpt_dir", str(tmp_path), "--mode", "decode", "--data_txt", str(tmp_path / "bundle.pt"),
                  "--batch_size", "2", "--num_inference_steps", "5", "--num_samples", "3", "--save", "--save_input",
                  "--save_conditioning", "--scheduler", "dpmsolver2", "--output_dir", str(tmp_path / "dec")])
    assert rc == 0
    dec = tmp_path / "dec" / "decode"
    assert torch.load(dec / "samples.pt", weights_only=True).shape == (3, 1, 32, 32)
    assert torch.equal(torch.load(dec / "input.pt", weights_only=True), tgt[:3])
    assert torch.equal(torch.load(dec / "conditioning.pt", weights_only=True), cond[:3])
    rc = RM.main(["--ckpt_dir", str(tmp_path), "--mode", "encode", "--targets_pt", str(tmp_path / "tgt.pt"),
                  "--timestep", "0", "--save", "--output_dir", str(tmp_path / "enc")])
    enc = torch.load(tmp_path / "enc" / "encode" / "encoded.pt", weights_only=True)
    assert rc == 0 and enc.shape == tgt.shape and float((enc - tgt).abs().max()) < 0.1   # abar_0 ~ 0.9999
