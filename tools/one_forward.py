"""Run N eager denoiser forwards of the headline config (for ncu captures). Usage: one_forward.py [B] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LDCT_UNET, synthetic_inputs  # noqa: E402
from fmdm_b200 import ops  # noqa: E402
from fmdm_b200.models.generators import DiffusionUNetFactory  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1).to(dev).eval()
noise, cond = synthetic_inputs(B, 42, dev)
t = torch.full((B,), 500.0, device=dev)
with torch.no_grad():
    for i in range(N):
        before = ops.launch_count()
        out = model(noise, t, context=cond)
        torch.cuda.synchronize()
        print(f"forward {i}: {ops.launch_count() - before} fmdm kernel launches, out mean {float(out.mean()):.5f}", flush=True)
