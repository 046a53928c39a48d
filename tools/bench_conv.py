"""Micro-benchmark of the conv implicit-GEMM kernel on the LDCT-512 shapes (SURVEY.md §8d shape list)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fmdm_b200 import ops  # noqa: E402

SHAPES = [  # (Cin segments, Cout, k, stride, H(in), count per forward)
    ([128], 128, 3, 1, 512, 8), ([128, 128], 128, 3, 1, 512, 3), ([128], 128, 3, 1, 256, 7),
    ([256], 256, 3, 1, 128, 7), ([256, 256], 256, 3, 1, 128, 2), ([256], 256, 3, 1, 256, 1),
    ([128, 128], 128, 3, 1, 256, 2), ([256, 128], 128, 3, 1, 256, 1), ([128, 128], 128, 1, 1, 512, 3),
    ([256], 256, 3, 1, 64, 7), ([512], 512, 3, 1, 32, 7), ([512], 512, 3, 1, 16, 11), ([512, 512], 512, 3, 1, 32, 2),
    ([128], 128, 3, 2, 512, 1), ([512], 1536, 1, 1, 32, 5),
    # ResBlock tail of the level-0 up blocks: conv2 3x3 (128 ch) + fused 1x1 skip conv over the (h, skip) concat
    ([128, 128, 128], 128, (3, 1, 1), 1, 512, 3), ([128, 128, 128], 128, (3, 1, 1), 1, 256, 3),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    tot_t = tot_f = 0.0
    sel = os.environ.get("SHAPES")
    shapes = SHAPES if not sel else [SHAPES[int(i)] for i in sel.split(",")]
    for cins, cout, k, s, H, cnt in shapes:
        xs = [torch.randn(B, H, H, c, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2) for c in cins]
        ks = list(k) if isinstance(k, (tuple, list)) else [k] * len(cins)
        ws = [torch.randn(cout, c, kk, kk, device=dev) * 0.05 for c, kk in zip(cins, ks)]
        pw = ops.pack_conv_weight([(w, 0, c) for w, c in zip(ws, cins)])
        bias = torch.randn(cout, device=dev)
        Ho = (H + s - 1) // s
        out = ops.empty_nhwc(B, cout, Ho, Ho, dev)
        norm = None
        if os.environ.get("NORM") and s == 1 and ks[0] == 3:
            tab = ops.NormTable(torch.rand(B, 2, sum(cins), device=dev) + 0.5, True)
            norm, off = [], 0
            for c, kk in zip(cins, ks):
                norm.append((tab, off) if kk == 3 else None)  # 1x1 skip segments read the raw input
                off += c
        res = None
        if os.environ.get("RESIDUAL") and s == 1:
            res = torch.randn(B, H, H, cout, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
        for _ in range(3):
            ops.conv2d(xs, pw, stride=s, bias=bias, out=out, norm=norm, residual=res)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d(xs, pw, stride=s, bias=bias, out=out, norm=norm, residual=res)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        flops = 2.0 * B * Ho * Ho * cout * sum(c * kk * kk for c, kk in zip(cins, ks))
        rows.append(dict(cin=cins, cout=cout, k=k, stride=s, H=H, ms=round(t, 4), tflops=round(flops / t / 1e9, 1)))
        tot_t += t * cnt
        tot_f += flops * cnt
        print(rows[-1], flush=True)
    print(json.dumps(dict(B=B, weighted_ms=tot_t, weighted_tflops=tot_f / tot_t / 1e9)))


if __name__ == "__main__":
    main()
