"""API conformance of the module mirror with the reference (`src.nn`, `src.models.unet`, factory, scheduler glue):
constructor kwargs, children / state_dict keys and shapes, seeded-init equality, `--scheduler` aliases."""
import glob
import inspect
import json
import os
import sys

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# denoiser manifests only; the AutoencoderKL decode-path manifests (state_keys_vae_*) are checked in test_vae_decode.py
CASES = sorted(n for n in (os.path.basename(p)[len("state_keys_"):-5]
                           for p in glob.glob(os.path.join(GOLD, "state_keys_*.json"))) if not n.startswith("vae_"))


@pytest.mark.parametrize("name", CASES)
def test_state_dict_keys_match_reference(name):
    from fmdm_b200.models.generators import DiffusionUNetFactory

    meta = json.load(open(os.path.join(GOLD, f"state_keys_{name}.json")))
    model = DiffusionUNetFactory().build(meta["cfg"], meta["conditioning"], 1)
    mine = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    assert mine == meta["keys"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the build container")
def test_seeded_init_and_signatures_match_live_reference():
    import importlib

    from fmdm_b200 import nn as my_nn
    from fmdm_b200.models import unet as my_unet
    from fmdm_b200.models.generators import DiffusionUNetFactory as Mine

    sys.path.insert(0, "/root/reference/src")
    try:
        ref_nn = importlib.import_module("nn")
        ref_blocks = importlib.import_module("nn.blocks")
        ref_att = importlib.import_module("nn.blocks.attention")
        ref_ops = importlib.import_module("nn.ops")

        def ref_cls(name):
            for mod in (ref_nn, ref_blocks, ref_att, ref_ops):
                if hasattr(mod, name):
                    return getattr(mod, name)
            raise AttributeError(name)

        ref_unet = importlib.import_module("models.unet")
        Ref = importlib.import_module("models.generators.diffusionfactory").DiffusionUNetFactory
        cfg = json.load(open("/root/reference/configs/MNIST/mnist_flow_matching_diffusers_nd.json"))["model"]["unet"]
        torch.manual_seed(0)
        a = Mine().build(cfg, "concatenate", 1).state_dict()
        torch.manual_seed(0)
        b = Ref().build(cfg, "concatenate", 1).state_dict()
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
        for cls in ("ConvND", "UpsampleND", "DownsampleND", "ResBlockND", "DiffusersAttentionND",
                    "SpatialSelfAttention", "SpatialCrossAttention", "QKVAttention", "DownBlock2DCompat",
                    "UpBlock2DCompat", "UNetMidBlock2DCompat", "PoolND", "UnPoolND", "AvgPoolND", "MaxPoolND",
                    "ConvTransposeND"):
            mine_sig = inspect.signature(getattr(my_nn, cls).__init__)
            ref_sig = inspect.signature(ref_cls(cls).__init__)
            assert list(mine_sig.parameters) == list(ref_sig.parameters), cls
            for p in ref_sig.parameters.values():
                assert mine_sig.parameters[p.name].default == p.default, (cls, p.name)
        for cls in ("UNetDiffusersND", "EfficientUNetND"):
            mine_sig = inspect.signature(getattr(my_unet, cls).__init__)
            ref_sig = inspect.signature(getattr(ref_unet, cls).__init__)
            assert list(mine_sig.parameters) == list(ref_sig.parameters), cls
        assert set(ref_unet.__all__) == set(my_unet.__all__)
        ref_public = {n for n in ref_nn.__all__ if n not in ("blocks", "ops", "losses", "PerceptualLoss",
                      "PatchDiscriminator", "discriminator_hinge_loss", "generator_hinge_loss", "vq_regularizer")}
        assert ref_public <= set(my_nn.__all__), ref_public - set(my_nn.__all__)
        assert set(ref_blocks.__all__) <= set(my_nn.__all__)
    finally:
        sys.path.remove("/root/reference/src")
        for k in list(sys.modules):
            if "/root/reference" in str(getattr(sys.modules[k], "__file__", "")):
                del sys.modules[k]


def test_factory_channel_arithmetic():
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.models.unet import EfficientUNetND, UNetDiffusersND

    f = DiffusionUNetFactory()
    m = f.build({"unet_impl": "diffusers_nd", "block_out_channels": [32, 64], "down_block_types": ["DownBlock2D"] * 2,
                 "up_block_types": ["UpBlock2D"] * 2}, "concatenate", 1)
    assert isinstance(m, UNetDiffusersND) and m.conv_in.in_channels == 2 and m.conv_out.out_channels == 1
    m = f.build({"unet_impl": "diffusers_nd", "block_out_channels": [32, 64], "down_block_types": ["DownBlock2D"] * 2,
                 "up_block_types": ["UpBlock2D"] * 2, "in_channels": 4, "out_channels": 4}, "concatenate", 4)
    assert m.conv_in.in_channels == 8 and m.conv_out.out_channels == 4
    m = f.build({"block_out_channels": [32, 32, 64], "attention_resolutions": []}, "concatenate", 1)
    assert isinstance(m, EfficientUNetND) and m.channel_mult == (1, 1, 2) and m.attention_resolutions == ()
    assert m.input_blocks[0][0].conv.in_channels == 2
    with pytest.raises(NotImplementedError):
        f.build({"block_out_channels": [32], "pool_factor": 2}, None, 1)


def test_scheduler_glue():
    from fmdm_b200.pipelines.schedulers import (DDIMScheduler, DPMSolverMultistepScheduler,
                                                FlowMatchEulerDiscreteScheduler)
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override

    s, n = build_scheduler({"name": "flow_match_euler", "num_train_timesteps": 1000, "num_inference_steps": 1000,
                            "params": {}}, {})
    assert isinstance(s, FlowMatchEulerDiscreteScheduler) and n == 1000 and s.config.num_train_timesteps == 1000
    assert not hasattr(s, "add_noise")  # the reference probes this with hasattr (diffusion_utils.py:218)
    ov = resolve_scheduler_override("dpmsolver++")
    assert ov == {"name": "dpm_multistep", "params": {"solver_order": 2, "algorithm_type": "dpmsolver++"}}
    spec = {"name": "ddpm", "params": {"beta_start": 1e-4, "beta_end": 0.02}, "num_inference_steps": 1000}
    merged = dict(spec, name=ov["name"], params={**spec["params"], **ov["params"]})
    s, _ = build_scheduler(merged, {"num_train_timesteps": 1000})
    assert isinstance(s, DPMSolverMultistepScheduler) and s.config.beta_end == 0.02
    s, _ = build_scheduler({"name": "ddim", "params": {"beta_start": 1e-4, "beta_end": 0.02, "bogus": 1}}, {})
    assert isinstance(s, DDIMScheduler) and hasattr(s, "add_noise")
    assert resolve_scheduler_override(None) is None and resolve_scheduler_override("  ") is None
    assert resolve_scheduler_override("FlowMatch") == {"name": "flow_match_euler"}
    with pytest.raises(ValueError):
        resolve_scheduler_override("nope")
    from fmdm_b200.pipelines.schedulers import DDPMScheduler

    s, n = build_scheduler({"name": "ddpm"}, {})   # the reference's default name (`pipelines/utils.py:46`)
    assert isinstance(s, DDPMScheduler) and n == 1000 and hasattr(s, "add_noise")
    s, _ = build_scheduler({}, {})
    assert isinstance(s, DDPMScheduler)
    from fmdm_b200.pipelines.schedulers import UniPCMultistepScheduler

    s, _ = build_scheduler({"name": "unipc"}, {})
    assert isinstance(s, UniPCMultistepScheduler)
    with pytest.raises(NotImplementedError):
        build_scheduler({"name": "dpm_sde"}, {})
    # dpmsolver1 / dpmsolver2: diffusers refuses algorithm_type "dpmsolver" with a zero final sigma; the aliases add the
    # `sigma_min` its error message asks for
    with pytest.raises(ValueError, match="sigma_min"):
        build_scheduler({"name": "dpm_multistep", "params": {"algorithm_type": "dpmsolver"}}, {})
    for alias, order in (("dpmsolver1", 1), ("dpmsolver2", 2)):
        ov = resolve_scheduler_override(alias)
        assert ov["params"]["algorithm_type"] == "dpmsolver" and ov["params"]["solver_order"] == order
        s, _ = build_scheduler({"name": ov["name"], "params": ov["params"]}, {})
        assert s.config.final_sigmas_type == "sigma_min" and s.config.solver_order == order
    with pytest.raises(ValueError):
        build_scheduler({"name": "nope"}, {})
    # training_cfg fallbacks
    s, n = build_scheduler({}, {"scheduler": "ddim", "num_train_timesteps": 500, "num_inference_steps": 25})
    assert isinstance(s, DDIMScheduler) and s.config.num_train_timesteps == 500 and n == 25


def test_product_scheduler_tables_match_oracle():
    """Host-side schedule construction of the product schedulers == the oracle's (timesteps, sigmas, per-step
    coefficients), so the GPU bit-exactness test only has the kernel arithmetic left to prove."""
    from fmdm_b200.pipelines.schedulers import (DDIMScheduler, DPMSolverMultistepScheduler,
                                                FlowMatchEulerDiscreteScheduler)
    from oracle.schedulers import DDIMOracle, DPMSolverPPOracle, FlowMatchEulerOracle

    for n in (1, 2, 20, 50, 1000):
        a, b = FlowMatchEulerDiscreteScheduler(1000), FlowMatchEulerOracle(1000)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
        assert torch.equal(a._coef_cpu[:, 0], b.sigmas[1:] - b.sigmas[:-1])
        assert a.plan_rows(a.timesteps[-3:]) == list(range(max(n - 3, 0), n))
    for n in (1, 10, 50):
        a, b = DDIMScheduler(1000, 1e-4, 0.02), DDIMOracle(1000, 1e-4, 0.02)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.alphas_cumprod, b.alphas_cumprod)
    from fmdm_b200.pipelines.schedulers import DDPMScheduler
    from oracle.schedulers import DDPMOracle

    for n in (1, 10, 50, 1000):
        a, b = DDPMScheduler(1000, 1e-4, 0.02), DDPMOracle(1000, 1e-4, 0.02)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.alphas_cumprod, b.alphas_cumprod)
        assert a._coef_cpu.shape == (n, 8) and float(a._coef_cpu[-1, 4]) == 0.0  # no noise on the last step (t = 0)
        if n > 1:
            t0 = int(b.timesteps[0])
            assert float(a._coef_cpu[0, 4]) == float(b._variance(t0, t0 - 1000 // n) ** 0.5)
    for n in (2, 5, 20):
        a, b = DPMSolverMultistepScheduler(1000, 1e-4, 0.02), DPMSolverPPOracle(1000, 1e-4, 0.02)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
        rows = a.plan_rows(a.timesteps)
        assert rows[0] == 0 and rows[-1] == n - 1 and all(r >= n for r in rows[1:-1])
        kw = dict(algorithm_type="dpmsolver", final_sigmas_type="sigma_min")
        a, b = DPMSolverMultistepScheduler(1000, 1e-4, 0.02, **kw), DPMSolverPPOracle(1000, 1e-4, 0.02, **kw)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas) and float(a.sigmas[-1]) > 0
        rows = a.plan_rows(a.timesteps)
        assert rows[0] == 0 and all(r >= n for r in rows[1:-1]) and (rows[-1] == n - 1) == (n < 15 or n == 1)
        assert torch.isfinite(a._coef_cpu).all()
    from fmdm_b200.pipelines.schedulers import UniPCMultistepScheduler
    from oracle.schedulers import UniPCOracle

    for n in (1, 2, 5, 20):
        a, b = UniPCMultistepScheduler(1000, 1e-4, 0.02), UniPCOracle(1000, 1e-4, 0.02)
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
        rows = a.plan_rows(a.timesteps)
        variants = [a._VARIANTS[r // n] for r in rows]
        want_pred = [1] + [2] * max(n - 2, 0) + ([1] if n > 1 else [])
        assert [v[0] for v in variants] == want_pred
        assert [v[1] for v in variants] == [0] + want_pred[:-1]
        assert torch.isfinite(a._coef_cpu[rows]).all()
        # a partial trajectory starts without a last sample: no corrector on its first step
        if n >= 5:
            sub = a.plan_rows(a.timesteps[2:])
            assert a._VARIANTS[sub[0] // n] == (1, 0) and a._VARIANTS[sub[1] // n] == (2, 1)
    with pytest.raises(ValueError):
        FlowMatchEulerDiscreteScheduler(1000).step(torch.zeros(1), 3, torch.zeros(1))
