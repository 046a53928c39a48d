"""Per-launch timing of one eager LDCT-512 denoiser forward (CUDA events around every fmdm kernel): tag, work, ms."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LDCT_UNET, synthetic_inputs  # noqa: E402
from fmdm_b200 import ops  # noqa: E402
from fmdm_b200.models.generators import DiffusionUNetFactory  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1).to(dev).eval()
noise, cond = synthetic_inputs(B, 42, dev)
t = torch.full((B,), 500.0, device=dev)
with torch.no_grad():
    model(noise, t, context=cond)
    model(noise, t, context=cond)
    with ops.profile() as rec:
        model(noise, t, context=cond)
tot = 0.0
for i, (tag, work, ms) in enumerate(rec.rows):
    tot += ms
    rate = work / ms / 1e9 if ms > 0 else 0.0
    unit = "TFLOP/s" if tag.startswith("conv") or tag == "attention" else "GB/s"
    if not unit.startswith("T"):
        rate *= 1e3 / 1e3
    print(f"{i:3d} {tag:18s} {ms * 1e3:9.1f} us  {rate:9.1f} {unit}")
print(f"total {tot:.3f} ms over {len(rec.rows)} timed launches")
