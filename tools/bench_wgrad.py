"""Micro-benchmark of the conv weight-gradient kernels on the LDCT-256 training shapes (B=16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fmdm_b200.training import functions as F  # noqa: E402

SHAPES = [(128, 128, 3, 1, 256), (256, 128, 3, 1, 256), (128, 128, 3, 1, 128), (256, 256, 3, 1, 64), (512, 256, 3, 1, 64),
          (256, 256, 3, 1, 32), (512, 512, 3, 1, 16), (512, 512, 3, 1, 8), (1024, 512, 3, 1, 16), (512, 1536, 1, 1, 16),
          (256, 128, 1, 1, 256), (128, 128, 3, 2, 256)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_t = tot_f = 0.0
    sel = os.environ.get("SHAPES")
    shapes = SHAPES if not sel else [SHAPES[int(i)] for i in sel.split(",")]
    for cin, cout, k, s, hw in shapes:
        x = torch.randn(B, hw, hw, cin, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        ho = hw // s
        dy = torch.randn(B, ho, ho, cout, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        dw = torch.empty((cout, cin, k, k) if k == 3 else (cout, cin), device=dev)
        for _ in range(2):
            F.conv_wgrad(dy, x, dw, ksize=k, stride=s, c_begin=0)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            F.conv_wgrad(dy, x, dw, ksize=k, stride=s, c_begin=0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[2]
        fl = 2.0 * B * ho * ho * cout * cin * k * k
        tot_t += t
        tot_f += fl
        print(f"cin={cin:5d} cout={cout:5d} k={k} s={s} hw={hw:4d}  {t:7.3f} ms  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
    print(f"total {tot_t:.3f} ms  {tot_f / tot_t / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
