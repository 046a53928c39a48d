"""ORACLE tooling: generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN MODULES
(imported read-only from /root/reference/src) on CPU fp32.  Run in the build container only:

    python oracle/make_golden.py

Fixtures (small, committed):
  denoiser_<name>.pt : {"cfg", "conditioning", "seed", "x", "t", "context", "out"} - reference model output for
                       weights = reinit_state_dict(reference_model.state_dict(), seed) (weights are re-derived in
                       the test from key names + shapes, so no state_dict is stored)
  state_keys_<name>.json : ordered [key, shape] list of the reference model's state_dict (API conformance)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from models.generators.diffusionfactory import DiffusionUNetFactory  # noqa: E402  (reference)

from oracle.denoiser import reinit_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    "mnist_diffusers_nd": dict(cfg_path="MNIST/mnist_flow_matching_diffusers_nd.json", cond="concatenate", hw=16, B=2),
    "ldct_diffusers_nd": dict(cfg_path="LDCT/LDCT_flow_matching_diffusers_nd.json", cond="concatenate", hw=32, B=1),
    "ldct_compvis": dict(cfg_path="LDCT/LDCT_flow_matching_compvis.json", cond="concatenate", hw=32, B=1),
    "ldct_ddpm_diffusers_nd": dict(cfg_path="LDCT/LDCT_ddpm_diffusers_nd.json", cond="concatenate", hw=32, B=1),
    "mnist_diffusers_nd_uncond": dict(cfg_path="MNIST/mnist_flow_matching_diffusers_nd.json", cond=None, hw=28, B=2),
}


# conditioning: "attention" (SURVEY.md §8f N4): cross-attention blocks over a latent context (B, 4, h, w)
CA_CASES = {
    "ca_diffusers_nd": dict(cfg={"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
                                 "block_out_channels": [64, 128], "cross_attention_dim": 4,
                                 "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D"],
                                 "mid_block_type": "UNetMidBlock2DCrossAttn",
                                 "up_block_types": ["CrossAttnUpBlock2D", "UpBlock2D"]}, hw=16, B=2, ctx_hw=8),
    "ca_efficient_nd": dict(cfg={"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1,
                                 "channel_mult": [1, 2], "model_channels": 64, "block_out_channels": [64, 128],
                                 "attention_resolutions": [2], "cross_attention_resolutions": [2],
                                 "cross_attention_in_middle": True, "cross_attention_dim": 4,
                                 "use_linear_attn": False}, hw=16, B=2, ctx_hw=8),
    # EfficientUNetND's default: linear attention inside the levels, softmax in the middle block
    "ca_efficient_nd_linear": dict(cfg={"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1,
                                        "num_res_blocks": 1, "channel_mult": [1, 2], "model_channels": 64,
                                        "block_out_channels": [64, 128], "attention_resolutions": [1, 2],
                                        "cross_attention_resolutions": [2], "cross_attention_in_middle": True,
                                        "cross_attention_dim": 4}, hw=16, B=2, ctx_hw=8),
}


def main_ca():
    for name, c in CA_CASES.items():
        cfg = c["cfg"]
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(cfg, "attention", 1).eval()
        keys = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        with open(os.path.join(GOLD, f"state_keys_{name}.json"), "w") as f:
            json.dump({"cfg": cfg, "conditioning": "attention", "keys": keys}, f)
        seed = 13
        model.load_state_dict(reinit_state_dict(model.state_dict(), seed))
        g = torch.Generator().manual_seed(77)
        x = torch.randn(c["B"], 1, c["hw"], c["hw"], generator=g)
        ctx = torch.randn(c["B"], 4, c["ctx_hw"], c["ctx_hw"], generator=g)
        t = torch.tensor([700.0, 12.0][: c["B"]])
        with torch.no_grad():
            out = model(x, t, context_ca=ctx)
        torch.save({"cfg": cfg, "conditioning": "attention", "seed": seed, "x": x, "t": t, "context": None,
                    "context_ca": ctx, "out": out}, os.path.join(GOLD, f"denoiser_{name}.pt"))
        print(name, tuple(out.shape), float(out.abs().max()), len(keys))


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, c in CASES.items():
        cfg = json.load(open(os.path.join("/root/reference/configs", c["cfg_path"])))["model"]["unet"]
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(cfg, c["cond"], 1).eval()
        keys = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        with open(os.path.join(GOLD, f"state_keys_{name}.json"), "w") as f:
            json.dump({"cfg": cfg, "conditioning": c["cond"], "keys": keys}, f)
        seed = 11
        sd = reinit_state_dict(model.state_dict(), seed)
        model.load_state_dict(sd)
        g = torch.Generator().manual_seed(123)
        x = torch.randn(c["B"], 1, c["hw"], c["hw"], generator=g)
        ctx = torch.rand(c["B"], 1, c["hw"], c["hw"], generator=g) if c["cond"] else None
        t = torch.tensor([979.6122, 21.3877][: c["B"]] if c["B"] > 1 else [500.5])
        with torch.no_grad():
            out = model(torch.cat([x, ctx], 1) if ctx is not None else x, t)
        torch.save({"cfg": cfg, "conditioning": c["cond"], "seed": seed, "x": x, "t": t, "context": ctx, "out": out},
                   os.path.join(GOLD, f"denoiser_{name}.pt"))
        print(name, tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    if "--ca-only" not in sys.argv:
        main()
    main_ca()
