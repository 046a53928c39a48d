"""Micro-benchmark of the GroupNorm(+SiLU) forward/backward streaming kernels on LDCT-256 training shapes (B=16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fmdm_b200.training import functions as F  # noqa: E402


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for c, hw in [(128, 256), (256, 256), (128, 128), (256, 64), (512, 16)]:
        B = 16
        x = torch.randn(B, hw, hw, c, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
        gamma = torch.rand(c, device=dev, requires_grad=True)
        beta = torch.rand(c, device=dev, requires_grad=True)
        gy = torch.randn(B, hw, hw, c, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        tf, tb = [], []
        for it in range(6):
            flush.zero_()
            e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e0.record()
            y = F.group_norm(x, gamma, beta, groups=32, eps=1e-5, silu=True)
            e1.record()
            flush.zero_()
            e1b = torch.cuda.Event(enable_timing=True)
            e1b.record()
            y.backward(gy)
            e2.record()
            torch.cuda.synchronize()
            if it:
                tf.append(e0.elapsed_time(e1))
                tb.append(e1b.elapsed_time(e2))
        nbytes = B * hw * hw * c * 2
        f, b = sorted(tf)[2], sorted(tb)[2]
        print(f"C={c:4d} hw={hw:4d}  fwd {f * 1e3:7.1f} us ({3 * nbytes / f / 1e9:6.2f} TB/s of 3 passes)   "
              f"bwd {b * 1e3:7.1f} us ({5 * nbytes / b / 1e9:6.2f} TB/s of 5 passes)", flush=True)


if __name__ == "__main__":
    main()
