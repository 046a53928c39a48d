"""Micro-benchmark of the non-GEMM kernels at the LDCT-512 shapes (B=16): stem, head (+fused norm), GroupNorm
table/finalize, upsample, attention.  CUDA events on the launching stream, L2 flushed between iterations."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fmdm_b200 import ops  # noqa: E402


def timeit(fn, flush, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    H = 512
    res = {}
    x0 = torch.randn(B, 1, H, H, device=dev)
    x1 = torch.rand(B, 1, H, H, device=dev)
    w = torch.randn(128, 2, 3, 3, device=dev) / 4
    b = torch.randn(128, device=dev)
    res["stem_ms"] = timeit(lambda: ops.conv_stem(x0, x1, w, b), flush)
    res["stem_nostats_ms"] = timeit(lambda: ops.conv_stem(x0, x1, w, b, want_stats=False), flush)
    res["stem_tc_ms"] = timeit(lambda: ops.conv_stem(x0, x1, w, b, tensor_cores=True), flush)
    res["stem_tc_nostats_ms"] = timeit(lambda: ops.conv_stem(x0, x1, w, b, want_stats=False, tensor_cores=True), flush)
    y = ops.conv_stem(x0, x1, w, b)
    gamma = torch.randn(128, device=dev)
    beta = torch.randn(128, device=dev)
    res["gn_table_from_stem_ms"] = timeit(lambda: ops.group_norm_table([y], 32, 1e-5, gamma, beta, silu=True), flush)
    pw = ops.pack_conv_weight([(torch.randn(128, 128, 3, 3, device=dev) * 0.03, 0, 128)])
    yc = ops.conv2d([y], pw, bias=b, want_stats=True)
    res["gn_table_from_conv_ms"] = timeit(lambda: ops.group_norm_table([yc], 32, 1e-5, gamma, beta, silu=True), flush)
    res["gn_table_concat_ms"] = timeit(
        lambda: ops.group_norm_table([yc, y], 32, 1e-5, torch.cat([gamma, gamma]), torch.cat([beta, beta]), silu=True),
        flush)
    tab = ops.group_norm_table([yc], 32, 1e-5, gamma, beta, silu=True)
    wh = torch.randn(1, 128, 3, 3, device=dev) / 30
    bh = torch.randn(1, device=dev)
    res["head_norm_ms"] = timeit(lambda: ops.conv_head(yc, wh, bh, norm=tab), flush)
    res["head_plain_ms"] = timeit(lambda: ops.conv_head(yc, wh, bh), flush)
    res["gn_apply_level0_ms"] = timeit(lambda: ops.group_norm([yc], 32, 1e-5, gamma, beta, silu=True), flush)
    xs = torch.randn(B, 256, 256, 128, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)
    res["upsample_256to512_c128_ms"] = timeit(lambda: ops.upsample_nearest2x(xs), flush)
    # attention as in the LDCT level-4 blocks: 64 heads x head_dim 8, T = 1024, qkv packed [B][T][3C]
    C_, T = 512, 1024
    qkv = torch.randn(B, T, 3 * C_, device=dev, dtype=torch.bfloat16)
    o = torch.empty(B, T, C_, device=dev, dtype=torch.bfloat16)
    res["attention_T1024_h64_d8_ms"] = timeit(
        lambda: ops.attention(qkv, qkv[:, :, C_:], qkv[:, :, 2 * C_:], o, batch=B, heads=64, tq=T, tk=T, head_dim=8,
                              q_strides=(T * 3 * C_, 8, 3 * C_), kv_strides=(T * 3 * C_, 8, 3 * C_),
                              o_strides=(T * C_, 8, C_)), flush)
    print(json.dumps({k: round(v, 4) for k, v in res.items()}))


if __name__ == "__main__":
    main()
