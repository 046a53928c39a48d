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
    assert RM.load_run_config(tmp_path)["model"]["model_type"] == "flow_matching"
    assert RM.resolve_checkpoint(tmp_path, "flow_matching") is None
    (tmp_path / "flow_last.pt").write_bytes(b"x")
    assert RM.resolve_checkpoint(tmp_path, "flow_matching").name == "flow_last.pt"
    (tmp_path / "flow_best.pt").write_bytes(b"x")
    assert RM.resolve_checkpoint(tmp_path, "flow_matching").name == "flow_best.pt"
    assert RM.resolve_checkpoint(tmp_path, "diffusion") is None


@pytest.mark.gpu
@pytest.mark.parametrize("sched,kw", [(None, {}), ("ddim", {"start_step": 500}), ("dpmsolver++", {"last_n_steps": 4}),
                                      ("flowmatch", {"num_inference_steps": 12})])
def test_decode_diffusion_batch_matches_oracle_loop(sched, kw):
    """decode_diffusion_batch == the oracle's restatement of the same call (scheduler override, subset, steps)."""
    from oracle import denoiser as OD
    from oracle.sampling import make_scheduler, sample_loop

    dev = torch.device("cuda")
    cfg = json.loads(json.dumps(CFG))
    cfg["model"]["unet"] = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
                            "block_out_channels": [64, 128], "down_block_types": ["DownBlock2D", "AttnDownBlock2D"],
                            "up_block_types": ["AttnUpBlock2D", "UpBlock2D"]}
    torch.manual_seed(5)
    model = DU.build_diffusion_model(cfg, dev)
    B, hw = 2, 32
    name = sched or "flowmatch"
    if name != "flowmatch":
        # epsilon samplers need a conditioned (trained) denoiser for a final-sample comparison: tests/_fixtures.py
        from _fixtures import train_epsilon_denoiser

        loss = train_epsilon_denoiser(model, hw=hw, batch=32, steps=300, lr=3e-4, warmup=50)
        assert loss < 0.1, f"the epsilon fixture did not train (loss {loss})"
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(B, 1, hw, hw, generator=g).to(dev)
    cond = torch.rand(B, 1, hw, hw, generator=g).to(dev)
    steps = kw.get("num_inference_steps", 20)
    with torch.no_grad():
        out = DU.decode_diffusion_batch(model, cfg["training"], cfg["model"], dev, tuple(noise.shape),
                                        conditioning_batch=cond, scheduler_override=sched,
                                        num_inference_steps=steps, start_step=kw.get("start_step"),
                                        last_n_steps=kw.get("last_n_steps"), init_sample=noise)

    def oracle_model(inp, t):
        return OD.unet_diffusers_nd_forward(sd, cfg["model"]["unet"], inp[:, :1], t, conditioning="concatenate",
                                            channels=1, context=inp[:, 1:])

    class _DevSched:  # oracle scheduler tables live on the CPU
        def __init__(self, s):
            self.s, self.set_timesteps = s, s.set_timesteps

        @property
        def timesteps(self):
            return self.s.timesteps

        def step(self, pred, t, x):
            r = self.s.step(pred.cpu(), t, x.cpu())
            r.prev_sample = r.prev_sample.to(dev)
            return r

    with torch.no_grad():
        ref = sample_loop(oracle_model, _DevSched(make_scheduler(name, 1000, {"beta_start": 1e-4, "beta_end": 0.02})),
                          steps, noise, cond, start_step=kw.get("start_step"), last_n_steps=kw.get("last_n_steps"))
    assert out.shape == ref.shape and torch.isfinite(out).all()
    mse = float(((out.clamp(0, 1) - ref.clamp(0, 1)) ** 2).mean())
    assert mse < 1e-4, (name, mse)  # north star: final samples >= 40 dB PSNR (peak 1.0 after clamp(0, 1))


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
