"""Diagnostic: per-step prediction error of the B200 denoiser vs the fp32 oracle under different inits."""
import sys, os, math, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import denoiser as OD
from fmdm_b200.models.generators import DiffusionUNetFactory
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
MNIST = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
         "block_out_channels": [64, 128, 128], "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
         "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
LDCT = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
        "block_out_channels": [128, 128, 256, 256, 512, 512],
        "down_block_types": ["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"],
        "up_block_types": ["UpBlock2D", "AttnUpBlock2D"] + ["UpBlock2D"] * 4}

def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())

for name, cfg, hw, B in [("mnist32", MNIST, 32, 4), ("ldct64", LDCT, 64, 2), ("ldct128", LDCT, 128, 1)]:
    for init in ("default", "reinit"):
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(cfg, "concatenate", 1)
        sd = model.state_dict()
        if init == "reinit":
            sd = OD.reinit_state_dict(sd, 1)
            model.load_state_dict(sd)
        model = model.to(DEV).eval()
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        sd_bf = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sdd.items()}
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
        c = torch.rand(B, 1, hw, hw, generator=g).to(DEV)
        errs, errs_w = [], []
        for tval in (999.0, 750.0, 500.5, 250.0, 1.0):
            t = torch.full((B,), tval, device=DEV)
            ref = OD.denoiser_forward(sdd, cfg, x, t, conditioning="concatenate", channels=1, context=c)
            refw = OD.denoiser_forward(sd_bf, cfg, x, t, conditioning="concatenate", channels=1, context=c)
            with torch.no_grad():
                out = model(x, t, context=c)
            errs.append(rel(out, ref)); errs_w.append(rel(refw, ref))
        print(name, init, "err", ["%.4f" % e for e in errs], "weight-only", ["%.4f" % e for e in errs_w], flush=True)
