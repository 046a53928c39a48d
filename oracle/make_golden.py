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
    main()
