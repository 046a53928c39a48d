"""Per-kernel parity on the B200: every C-ABI kernel against a plain PyTorch fp32 reference of the same op
(the floating-point kernels) or against the oracle bit-for-bit (the scheduler steps)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from fmdm_b200 import ops  # noqa: E402

DEV = "cuda"
# the fp32 references must be true fp32 (cuDNN/cuBLAS default to TF32 for convs)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _nhwc(t):
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def _conv_case(B, H, W, cins, cout, ks_list, stride=1, bias=True, addvec=False, residual=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [_bf16r(torch.randn(B, c, H, W, generator=g)).to(DEV) for c in cins]
    ws = [_bf16r(torch.randn(cout, c, k, k, generator=g) / math.sqrt(c * k * k)).to(DEV) for c, k in zip(cins, ks_list)]
    bvec = torch.randn(cout, generator=g).to(DEV) if bias else None
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    av = torch.randn(B, cout, generator=g).to(DEV) if addvec else None
    res = _bf16r(torch.randn(B, cout, Ho, Wo, generator=g)).to(DEV) if residual else None
    ref = torch.zeros(B, cout, Ho, Wo, device=DEV)
    for x, w, k in zip(xs, ws, ks_list):
        ref = ref + F.conv2d(x, w, None, stride=stride, padding=k // 2)
    if bias:
        ref = ref + bvec.view(1, -1, 1, 1)
    if addvec:
        ref = ref + av.view(B, cout, 1, 1)
    if residual:
        ref = ref + res
    packed = ops.pack_conv_weight([(w, 0, c) for w, c in zip(ws, cins)])
    out = ops.conv2d([_nhwc(x) for x in xs], packed, stride=stride, bias=bvec, addvec=av,
                     residual=_nhwc(res) if residual else None)
    torch.cuda.synchronize()
    return out.float(), ref


CONV_CASES = [
    # B, H, W, cins, cout, ks, stride, bias, addvec, residual
    (1, 16, 16, [64], 64, [1], 1, False, False, False),
    (1, 16, 16, [128], 128, [1], 1, True, False, False),
    (1, 16, 16, [64], 128, [3], 1, True, False, False),
    (2, 32, 32, [128], 256, [3], 1, True, True, True),
    (2, 32, 32, [128], 128, [3], 2, True, False, False),
    (2, 16, 16, [64, 128], 128, [3, 3], 1, True, True, False),
    (2, 16, 16, [128, 64, 128], 128, [3, 1, 1], 1, True, True, False),
    (3, 28, 28, [64], 64, [3], 1, True, False, True),
    (3, 14, 14, [128], 128, [3], 1, True, True, True),
    (3, 7, 7, [128], 128, [3], 1, True, True, True),
    (3, 28, 28, [64], 64, [3], 2, True, False, False),
    (3, 14, 14, [128], 128, [3], 2, True, False, False),
    (2, 16, 16, [1024], 512, [3], 1, True, True, True),
    (1, 16, 16, [32], 72, [3], 1, True, False, False),
    (1, 64, 256, [128], 128, [3], 1, True, True, True),
    (2, 33, 35, [64], 64, [3], 2, True, False, False),
    (1, 1024, 1, [512], 1536, [1], 1, True, False, False),
    # row mode (Wt == 128): kw taps served from one 130-pixel halo row
    (1, 8, 128, [64], 64, [3], 1, True, False, False),
    (2, 5, 256, [128], 128, [3], 1, True, True, True),
    (1, 6, 200, [64, 128], 256, [3, 3], 1, True, True, False),
    (2, 4, 130, [128, 64, 64], 128, [3, 1, 1], 1, True, False, False),
    (1, 3, 512, [256], 512, [3], 1, True, False, True),
    (1, 7, 100, [32], 72, [3], 1, True, False, False),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "B{}_{}x{}_cin{}_cout{}_k{}_s{}".format(
    c[0], c[1], c[2], "+".join(map(str, c[3])), c[4], "".join(map(str, c[5])), c[6]))
def test_conv2d_igemm(case):
    out, ref = _conv_case(*case)
    err = _rel_l2(out, ref)
    assert err < 6e-3, f"rel L2 {err}"
    assert float((out - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 0.05


@pytest.mark.parametrize("case", [
    (2, 5, 256, [128], 128, [3], 1, True, True, True),          # row mode, ragged unit count (10 tiles -> 3 units)
    (3, 28, 28, [64], 128, [3], 1, True, True, True),           # box mode, tiles straddle nothing, odd tile count
    (2, 4, 130, [128, 64, 64], 128, [3, 1, 1], 1, True, False, False),
    (3, 33, 35, [64], 96, [3], 2, True, False, False),
    (2, 512, 512, [128], 128, [3], 1, True, True, True),        # the LDCT level-0 shape (picked without the override)
], ids=["row", "box", "row_3seg", "box_s2", "ldct_level0"])
def test_conv2d_igemm_two_m_tiles_per_cta(case, monkeypatch):
    """The MT = 2 variant (two M tiles per CTA share every weight tile) against the same fp32 reference."""
    monkeypatch.setenv("FMDM_CONV_MT", "2")
    out, ref = _conv_case(*case)
    err = _rel_l2(out, ref)
    assert err < 6e-3, f"rel L2 {err}"
    assert float((out - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 0.05


@pytest.mark.parametrize("case", [
    (3, 28, 28, [64], 64, [3], 1), (2, 14, 14, [128], 128, [3], 1), (3, 7, 7, [128, 64], 128, [3, 1], 1),
    (2, 28, 28, [64], 128, [3], 2), (1, 16, 16, [32], 72, [1], 1), (2, 40, 36, [64, 64], 64, [3, 3], 1),
], ids=lambda c: "B{}_{}x{}_cin{}_cout{}_k{}_s{}".format(c[0], c[1], c[2], "+".join(map(str, c[3])), c[4],
                                                        "".join(map(str, c[5])), c[6]))
def test_conv2d_split_weights_share_the_a_tile(case, monkeypatch):
    """Split-bf16 weights (w_hi + w_lo along K) on the per-tile kernel: each A tile is multiplied by its hi and its lo
    weight tile (`dup_koff`) instead of being fetched once per half.  Equal to the two-pass form (FMDM_CONV_NO_DUP=1)
    up to the order of the fp32 accumulation, and much closer to the fp32-weight reference than plain bf16 weights."""
    B, H, W, cins, cout, ks, stride = case
    g = torch.Generator(device="cpu").manual_seed(9)
    xs = [_bf16r(torch.randn(B, c, H, W, generator=g)).to(DEV) for c in cins]
    ws = [(torch.randn(cout, c, k, k, generator=g) / math.sqrt(c * k * k)).to(DEV) for c, k in zip(cins, ks)]
    bvec = torch.randn(cout, generator=g).to(DEV)
    ref = bvec.view(1, -1, 1, 1)
    for x, w, k in zip(xs, ws, ks):
        ref = ref + F.conv2d(x, w, None, stride=stride, padding=k // 2)
    srcs = [_nhwc(x) for x in xs]
    parts = [(w, 0, c) for w, c in zip(ws, cins)]
    split = ops.pack_conv_weight(parts, split=True)
    plain = ops.pack_conv_weight(parts)
    assert split.split
    out = ops.conv2d(srcs, split, stride=stride, bias=bvec).float()
    monkeypatch.setenv("FMDM_CONV_NO_DUP", "1")
    two_pass = ops.conv2d(srcs, split, stride=stride, bias=bvec).float()
    monkeypatch.delenv("FMDM_CONV_NO_DUP")
    out_plain = ops.conv2d(srcs, plain, stride=stride, bias=bvec).float()
    assert _rel_l2(out, two_pass) < 2e-3                      # both carry the bf16 rounding of the OUTPUT only
    assert _rel_l2(out, ref) < 3e-3                           # output rounding (2^-9) dominates
    # against the un-rounded output the weight error is gone: compare in fp32 through the rounding of the reference
    assert _rel_l2(out, _bf16r(ref)) < 0.7 * _rel_l2(out_plain, _bf16r(ref)) + 2e-4


ROLLING_CASES = [c for c in CONV_CASES if c[6] == 1 and c[2] > 64 and c[5][0] == 3] + [
    (2, 37, 128, [64], 64, [3], 1, True, True, True),
    (3, 70, 200, [128, 64], 128, [3, 3], 1, True, True, False),
    (1, 33, 512, [128, 128, 128], 256, [3, 1, 1], 1, True, False, False),
    (5, 1, 128, [64], 128, [3], 1, True, False, True),          # single-row images, odd strip count
    (1, 2, 130, [64], 192, [3], 1, False, False, False),
]


@pytest.mark.parametrize("mode", ["0", "3"], ids=["per_tile", "rolling"])
@pytest.mark.parametrize("case", ROLLING_CASES, ids=lambda c: "B{}_{}x{}_cin{}_cout{}_k{}".format(
    c[0], c[1], c[2], "+".join(map(str, c[3])), c[4], "".join(map(str, c[5]))))
def test_conv2d_igemm_rolling_rows(case, mode, monkeypatch):
    """Row-mode shapes through both schedules: per-tile K loop and the rolling-row strips (forced for every Cout)."""
    monkeypatch.setenv("FMDM_CONV_ROLLING", mode)
    out, ref = _conv_case(*case)
    err = _rel_l2(out, ref)
    assert err < 6e-3, f"rel L2 {err}"
    assert float((out - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 0.05


def test_group_norm_silu():
    g = torch.Generator().manual_seed(1)
    for (B, C, H, W, groups) in [(2, 128, 32, 32, 32), (3, 64, 7, 7, 32), (2, 512, 16, 16, 32), (1, 256, 64, 64, 32)]:
        x = _bf16r(torch.randn(B, C, H, W, generator=g) * 2 + 0.5).to(DEV)
        gamma = torch.randn(C, generator=g).to(DEV)
        beta = torch.randn(C, generator=g).to(DEV)
        for silu in (True, False):
            ref = F.group_norm(x, groups, gamma, beta, 1e-5)
            if silu:
                ref = F.silu(ref)
            out = ops.group_norm([_nhwc(x)], groups, 1e-5, gamma, beta, silu=silu).float()
            assert _rel_l2(out, ref) < 5e-3


def test_group_norm_concat_and_scale_shift():
    g = torch.Generator().manual_seed(2)
    B, C0, C1, H, W = 2, 256, 128, 16, 16
    x0 = _bf16r(torch.randn(B, C0, H, W, generator=g)).to(DEV)
    x1 = _bf16r(torch.randn(B, C1, H, W, generator=g) * 3 - 1).to(DEV)
    gamma = torch.randn(C0 + C1, generator=g).to(DEV)
    beta = torch.randn(C0 + C1, generator=g).to(DEV)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], 1), 32, gamma, beta, 1e-5))
    out = ops.group_norm([_nhwc(x0), _nhwc(x1)], 32, 1e-5, gamma, beta, silu=True).float()
    assert _rel_l2(out, ref) < 5e-3
    ss = torch.randn(B, 2 * C0, generator=g).to(DEV) * 0.5
    gam0, bet0 = gamma[:C0].contiguous(), beta[:C0].contiguous()
    ref = F.group_norm(x0, 32, gam0, bet0, 1e-5) * (1 + ss[:, :C0, None, None]) + ss[:, C0:, None, None]
    ref = F.silu(ref)
    out = ops.group_norm([_nhwc(x0)], 32, 1e-5, gam0, bet0, silu=True, scale_shift=ss).float()
    assert _rel_l2(out, ref) < 5e-3


def test_stem_and_head():
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 40, 36
    x0 = torch.randn(B, 1, H, W, generator=g).to(DEV)
    x1 = torch.rand(B, 1, H, W, generator=g).to(DEV)
    w = (torch.randn(128, 2, 3, 3, generator=g) / 4).to(DEV)
    b = torch.randn(128, generator=g).to(DEV)
    ref = F.conv2d(torch.cat([x0, x1], 1), w, b, padding=1)
    out = ops.conv_stem(x0, x1, w, b).float()
    assert _rel_l2(out, ref) < 4e-3
    ref = F.conv2d(2 * x0 - 1, w[:, :1].contiguous(), b, padding=1)
    out = ops.conv_stem(x0, None, w[:, :1].contiguous(), b, in_scale=2.0, in_shift=-1.0).float()
    assert _rel_l2(out, ref) < 4e-3
    for cout in (1, 4):
        xh = _bf16r(torch.randn(B, 128, H, W, generator=g)).to(DEV)
        wh = (torch.randn(cout, 128, 3, 3, generator=g) / 30).to(DEV)
        bh = torch.randn(cout, generator=g).to(DEV)
        ref = F.conv2d(xh, wh, bh, padding=1)
        out = ops.conv_head(_nhwc(xh), wh, bh)
        assert out.dtype == torch.float32 and out.is_contiguous()
        assert _rel_l2(out, ref) < 1e-4


def test_upsample_transpose():
    g = torch.Generator().manual_seed(4)
    x = _bf16r(torch.randn(2, 64, 9, 7, generator=g)).to(DEV)
    out = ops.upsample_nearest2x(_nhwc(x)).float()
    assert torch.equal(out, F.interpolate(x, scale_factor=2, mode="nearest"))
    t = torch.randn(3, 50, 72, generator=g).to(DEV).to(torch.bfloat16)
    assert torch.equal(ops.transpose_bf16(t), t.transpose(1, 2).contiguous())


@pytest.mark.parametrize("hd,heads,T", [(8, 64, 1024), (8, 16, 256), (64, 4, 256), (16, 8, 100), (32, 2, 77),
                                        (64, 4, 4096), (64, 1, 130), (32, 3, 1000), (16, 2, 64)])
def test_attention(hd, heads, T):
    g = torch.Generator().manual_seed(5)
    B = 2
    C = hd * heads
    qkv = _bf16r(torch.randn(B, T, 3 * C, generator=g)).to(DEV)
    q, k, v = [t.reshape(B, T, heads, hd).transpose(1, 2) for t in qkv.split(C, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, C)
    qkv16 = qkv.to(torch.bfloat16).contiguous()
    out = torch.empty(B, T, C, dtype=torch.bfloat16, device=DEV)
    ops.attention(qkv16, qkv16[:, :, C:], qkv16[:, :, 2 * C:], out, batch=B, heads=heads, tq=T, tk=T, head_dim=hd,
                  q_strides=(T * 3 * C, hd, 3 * C), kv_strides=(T * 3 * C, hd, 3 * C), o_strides=(T * C, hd, C))
    assert _rel_l2(out.float(), ref) < 6e-3


def test_timestep_embedding_and_linear():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.denoiser import timestep_embedding as ref_temb

    t = torch.tensor([1000.0, 979.6122, 500.5, 21.3877, 1.0, 0.0], device=DEV)
    for dim, flip, shift in [(128, True, 0), (64, False, 0), (128, True, 1), (33, False, 0)]:
        out = ops.timestep_embedding(t, dim, flip_sin_to_cos=flip, freq_shift=shift)
        ref = ref_temb(t.cpu(), dim, flip_sin_to_cos=flip, freq_shift=shift).to(DEV)
        assert float((out - ref).abs().max()) < 2e-4
    g = torch.Generator().manual_seed(6)
    x = torch.randn(16, 512, generator=g).to(DEV)
    w = (torch.randn(384, 512, generator=g) / 22).to(DEV)
    b = torch.randn(384, generator=g).to(DEV)
    b2 = torch.randn(384, generator=g).to(DEV)
    ref = F.linear(F.silu(x), w, b) + b2
    out = ops.linear_f32(x, w, b, b2, silu_in=True)
    assert float((out - ref).abs().max()) < 1e-4
    ref = F.silu(F.linear(x, w, b))
    out = ops.linear_f32(x, w, b, silu_out=True)
    assert float((out - ref).abs().max()) < 1e-4


def test_conv_fused_groupnorm_statistics():
    """GroupNorm fed by the conv epilogue's channel-quad partial sums == GroupNorm with its own statistics pass."""
    g = torch.Generator().manual_seed(11)
    for (B, H, W, cin, cout, stride) in [(2, 32, 32, 64, 128, 1), (3, 28, 28, 64, 64, 1), (2, 32, 32, 128, 256, 2),
                                         (1, 16, 48, 64, 512, 1),
                                         # tiny images: two / four whole images per 128-row M tile (ragged last tile)
                                         (5, 7, 7, 128, 128, 1), (3, 8, 8, 64, 256, 1), (7, 4, 8, 64, 64, 1),
                                         (6, 14, 14, 128, 128, 2)]:
        x = _bf16r(torch.randn(B, cin, H, W, generator=g)).to(DEV)
        w = _bf16r(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(DEV)
        bias = torch.randn(cout, generator=g).to(DEV)
        res = None
        if stride == 1:
            res = _bf16r(torch.randn(B, cout, H, W, generator=g)).to(DEV)
        pw = ops.pack_conv_weight([(w, 0, cin)])
        y = ops.conv2d([_nhwc(x)], pw, stride=stride, bias=bias, residual=None if res is None else _nhwc(res),
                       want_stats=True)
        assert hasattr(y, "_fm_stats")
        gamma = torch.randn(cout, generator=g).to(DEV)
        beta = torch.randn(cout, generator=g).to(DEV)
        fused = ops.group_norm([y], 32, 1e-5, gamma, beta, silu=True).float()
        plain = ops.group_norm([y.clone(memory_format=torch.preserve_format)], 32, 1e-5, gamma, beta, silu=True).float()
        ref = F.silu(F.group_norm(y.float(), 32, gamma, beta, 1e-5))
        assert _rel_l2(fused, ref) < 5e-3 and _rel_l2(fused, plain) < 2e-3
    # the two-M-tiles-per-CTA variant writes the same partial-sum layout (B=2, 256x256 -> 1024 tiles; forced)
    import os
    os.environ["FMDM_CONV_MT"] = "2"
    try:
        x = _bf16r(torch.randn(2, 64, 64, 256, generator=g)).to(DEV)
        w = _bf16r(torch.randn(128, 64, 3, 3, generator=g) / 24).to(DEV)
        y = ops.conv2d([_nhwc(x)], ops.pack_conv_weight([(w, 0, 64)]), want_stats=True)
        gamma = torch.randn(128, generator=g).to(DEV)
        beta = torch.randn(128, generator=g).to(DEV)
        fused = ops.group_norm([y], 32, 1e-5, gamma, beta, silu=True).float()
        ref = F.silu(F.group_norm(y.float(), 32, gamma, beta, 1e-5))
        assert _rel_l2(fused, ref) < 5e-3
    finally:
        del os.environ["FMDM_CONV_MT"]
    # virtual concat of two producers with a group size that straddles neither source evenly (384 ch -> 12/group)
    xa = _bf16r(torch.randn(2, 64, 16, 16, generator=g)).to(DEV)
    wa = _bf16r(torch.randn(256, 64, 3, 3, generator=g) / 24).to(DEV)
    wb = _bf16r(torch.randn(128, 64, 3, 3, generator=g) / 24).to(DEV)
    ya = ops.conv2d([_nhwc(xa)], ops.pack_conv_weight([(wa, 0, 64)]), want_stats=True)
    yb = ops.conv2d([_nhwc(xa)], ops.pack_conv_weight([(wb, 0, 64)]), want_stats=True)
    gamma = torch.randn(384, generator=g).to(DEV)
    beta = torch.randn(384, generator=g).to(DEV)
    fused = ops.group_norm([ya, yb], 32, 1e-5, gamma, beta, silu=False).float()
    ref = F.group_norm(torch.cat([ya.float(), yb.float()], 1), 32, gamma, beta, 1e-5)
    assert _rel_l2(fused, ref) < 5e-3


# ---------------------------------------------------------------------------------------------------------------
# GroupNorm apply + SiLU folded into the conv operand path (XF kernels) and into the head conv's load
# ---------------------------------------------------------------------------------------------------------------
def _norm_conv_case(B, H, W, cins, normed, cout, ks_list, silu=True, residual=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [_bf16r(torch.randn(B, c, H, W, generator=g) * 1.5 + 0.3).to(DEV) for c in cins]
    ws = [_bf16r(torch.randn(cout, c, k, k, generator=g) / math.sqrt(c * k * k)).to(DEV) for c, k in zip(cins, ks_list)]
    bvec = torch.randn(cout, generator=g).to(DEV)
    ctot = sum(c for c, n in zip(cins, normed) if n)
    ab = torch.empty(B, 2, ctot)
    ab[:, 0] = torch.rand(B, ctot, generator=g) + 0.5
    ab[:, 1] = torch.randn(B, ctot, generator=g) * 0.5
    ab = ab.to(DEV)
    table = ops.NormTable(ab, silu)
    res = _bf16r(torch.randn(B, cout, H, W, generator=g)).to(DEV) if residual else None
    ref = torch.zeros(B, cout, H, W, device=DEV)
    norm, off = [], 0
    for x, w, k, c, n in zip(xs, ws, ks_list, cins, normed):
        if n:
            y = x * ab[:, 0, off:off + c, None, None] + ab[:, 1, off:off + c, None, None]
            y = _bf16r(F.silu(y) if silu else y)  # the operand reaches the tensor core in bf16
            norm.append((table, off))
            off += c
        else:
            y = x
            norm.append(None)
        ref = ref + F.conv2d(y, w, None, padding=k // 2)
    ref = ref + bvec.view(1, -1, 1, 1)
    if residual:
        ref = ref + res
    packed = ops.pack_conv_weight([(w, 0, c) for w, c in zip(ws, cins)])
    out = ops.conv2d([_nhwc(x) for x in xs], packed, bias=bvec, residual=_nhwc(res) if residual else None, norm=norm)
    torch.cuda.synchronize()
    return out.float(), ref


NORM_CONV_CASES = [
    # B, H, W, cins, normed, cout, ks, silu, residual
    (2, 5, 256, [128], [True], 128, [3], True, True),
    (1, 6, 200, [64, 128], [True, True], 256, [3, 3], True, False),        # concat GN1 of an up block, ragged W
    (2, 4, 130, [128, 64, 64], [True, False, False], 128, [3, 1, 1], True, False),  # conv2 + fused 1x1 skip (raw)
    (1, 3, 512, [256], [True], 512, [3], True, True),
    (1, 8, 128, [64], [True], 64, [3], False, False),                      # no activation
    (3, 9, 128, [64], [True], 64, [3], True, False),                       # odd tile count
    (2, 64, 128, [128], [True], 128, [3], True, True),
]


@pytest.mark.parametrize("case", NORM_CONV_CASES, ids=lambda c: "B{}_{}x{}_cin{}_cout{}".format(
    c[0], c[1], c[2], "+".join(map(str, c[3])), c[5]))
def test_conv2d_operand_norm(case):
    """conv(act(a*x+b)) with the transform running inside the conv kernel == the same conv on a pre-activated input."""
    out, ref = _norm_conv_case(*case)
    err = _rel_l2(out, ref)
    assert err < 6e-3, f"rel L2 {err}"
    assert float((out - ref).abs().max()) < 0.05 * float(ref.abs().max()) + 0.05


def test_conv2d_operand_norm_level0_shape():
    out, ref = _norm_conv_case(2, 512, 512, [128, 128], [True, True], 128, [3, 3], True, False, seed=3)
    assert _rel_l2(out, ref) < 6e-3


def test_group_norm_table_and_head():
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 128, 24, 40
    x = _bf16r(torch.randn(B, C, H, W, generator=g) * 2 + 0.5).to(DEV)
    gamma = torch.randn(C, generator=g).to(DEV)
    beta = torch.randn(C, generator=g).to(DEV)
    ss = torch.randn(B, 2 * C, generator=g).to(DEV) * 0.3
    tab = ops.group_norm_table([_nhwc(x)], 32, 1e-5, gamma, beta, silu=True, scale_shift=ss)
    y = x * tab.ab[:, 0, :, None, None] + tab.ab[:, 1, :, None, None]
    ref = F.group_norm(x, 32, gamma, beta, 1e-5) * (1 + ss[:, :C, None, None]) + ss[:, C:, None, None]
    assert float((y - ref).abs().max()) < 2e-3
    # table from conv-epilogue partial statistics == table from the statistics pass
    w = _bf16r(torch.randn(C, 64, 3, 3, generator=g) / 24).to(DEV)
    xin = _bf16r(torch.randn(B, 64, H, W, generator=g)).to(DEV)
    yc = ops.conv2d([_nhwc(xin)], ops.pack_conv_weight([(w, 0, 64)]), want_stats=True)
    t_fused = ops.group_norm_table([yc], 32, 1e-5, gamma, beta, silu=True)
    t_plain = ops.group_norm_table([yc.clone(memory_format=torch.preserve_format)], 32, 1e-5, gamma, beta, silu=True)
    assert float((t_fused.ab - t_plain.ab).abs().max()) < 2e-3
    # head conv with the output norm folded into its load
    wh = torch.randn(1, C, 3, 3, generator=g).to(DEV) / 30
    bh = torch.randn(1, generator=g).to(DEV)
    tab = ops.group_norm_table([_nhwc(x)], 32, 1e-5, gamma, beta, silu=True)
    out = ops.conv_head(_nhwc(x), wh, bh, norm=tab)
    ref = F.conv2d(F.silu(F.group_norm(x, 32, gamma, beta, 1e-5)), wh, bh, padding=1)
    assert _rel_l2(out, ref) < 5e-3


@pytest.mark.parametrize("cin,H,W,norm", [(128, 40, 36, False), (64, 17, 130, True), (128, 33, 200, True),
                                          (128, 16, 64, True), (64, 1, 5, False)])
def test_conv_head_dot_kernel(cin, H, W, norm):
    """Cout = 1 head (dot-then-gather kernel) with and without the fused output norm, ragged tiles included."""
    g = torch.Generator().manual_seed(7)
    B = 3
    x = _bf16r(torch.randn(B, cin, H, W, generator=g) * 1.3 + 0.2).to(DEV)
    wh = (torch.randn(1, cin, 3, 3, generator=g) / 20).to(DEV)
    bh = torch.randn(1, generator=g).to(DEV)
    tab = None
    y = x
    if norm:
        ab = torch.empty(B, 2, cin)
        ab[:, 0] = torch.rand(B, cin, generator=g) + 0.5
        ab[:, 1] = torch.randn(B, cin, generator=g) * 0.5
        ab = ab.to(DEV)
        tab = ops.NormTable(ab, True)
        y = F.silu(x * ab[:, 0, :, None, None] + ab[:, 1, :, None, None])
    ref = F.conv2d(y, wh, bh, padding=1)
    out = ops.conv_head(_nhwc(x), wh, bh, norm=tab)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel_l2(out, ref) < (2e-3 if norm else 1e-4)


@pytest.mark.parametrize("cin,H,W,silu,wscale", [(128, 130, 70, True, 1 / 20), (64, 200, 129, True, 1 / 20),
                                                 (128, 128, 64, False, 1 / 20), (128, 33, 200, True, 3e-7),
                                                 (64, 17, 130, False, 4e3), (128, 160, 66, True, 0.0)])
def test_conv_head_tensor_core_kernel(cin, H, W, silu, wscale):
    """Cout = 1 head with the fused output norm: phase 1 on `mma.sync` (fp16 operands, power-of-two weight scaling).
    Both tile heights, ragged tiles, with / without SiLU, weights far outside fp16's range, and an all-zero conv_out
    (the reference zero-initialises it)."""
    g = torch.Generator().manual_seed(71)
    B = 2
    x = _bf16r(torch.randn(B, cin, H, W, generator=g) * 1.3 + 0.2).to(DEV)
    wh = (torch.randn(1, cin, 3, 3, generator=g) * wscale).to(DEV)
    bh = (torch.randn(1, generator=g) * max(wscale, 1e-9) * 20).to(DEV)
    ab = torch.empty(B, 2, cin)
    ab[:, 0] = torch.rand(B, cin, generator=g) + 0.5
    ab[:, 1] = torch.randn(B, cin, generator=g) * 0.5
    ab = ab.to(DEV)
    y = x * ab[:, 0, :, None, None] + ab[:, 1, :, None, None]
    ref = F.conv2d(F.silu(y) if silu else y, wh, bh, padding=1)
    out = ops.conv_head(_nhwc(x), wh, bh, norm=ops.NormTable(ab, silu))
    assert out.shape == ref.shape and out.dtype == torch.float32
    if wscale == 0.0:
        assert torch.equal(out, ref)
    else:
        assert _rel_l2(out, ref) < 2e-3


def test_stem_fused_groupnorm_statistics():
    """GroupNorm table from the stem kernel's partial statistics == table from a statistics pass over its output."""
    g = torch.Generator().manual_seed(12)
    for (B, H, W, cout) in [(2, 40, 36, 128), (3, 17, 130, 64), (1, 128, 128, 128)]:
        x0 = torch.randn(B, 1, H, W, generator=g).to(DEV)
        x1 = torch.rand(B, 1, H, W, generator=g).to(DEV)
        w = (torch.randn(cout, 2, 3, 3, generator=g) / 4).to(DEV)
        b = torch.randn(cout, generator=g).to(DEV)
        y = ops.conv_stem(x0, x1, w, b)
        assert hasattr(y, "_fm_stats")
        gamma = torch.randn(cout, generator=g).to(DEV)
        beta = torch.randn(cout, generator=g).to(DEV)
        t_fused = ops.group_norm_table([y], 32, 1e-5, gamma, beta, silu=True)
        t_plain = ops.group_norm_table([y.clone(memory_format=torch.preserve_format)], 32, 1e-5, gamma, beta, silu=True)
        # the fused statistics see the fp32 accumulators, the plain pass their bf16 rounding
        assert float((t_fused.ab - t_plain.ab).abs().max()) < 5e-3 * float(t_plain.ab.abs().max())


@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 512, 512, 2, 128), (6, 256, 200, 1, 128), (2, 300, 444, 4, 64),
                                             (1, 512, 520, 8, 256), (3, 300, 444, 3, 64), (2, 515, 300, 2, 64),
                                             (17, 128, 136, 2, 128)])
def test_stem_tensor_core_path(B, H, W, cin, cout):
    """conv_in on large inputs: ONE `mma.sync` launch (`fm_conv_stem_tc_f32_bf16`: Cout 64 / 128, Cin <= 3) or
    `fm_stem_im2col_bf16` + the 1x1 implicit GEMM: against F.conv2d on the bf16-rounded
    operands (tight) and on the fp32 operands (bf16 rounding of inputs and weights only), with the 2x-1 centering, the
    fused conditioning concat and the GroupNorm partial statistics of the output."""
    g = torch.Generator().manual_seed(17)
    c0 = max(1, cin // 2)
    x0 = torch.randn(B, c0, H, W, generator=g).to(DEV)
    x1 = torch.rand(B, cin - c0, H, W, generator=g).to(DEV) if cin > c0 else None
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 4).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    assert B * H * W >= ops.STEM_TENSOR_MIN_PIXELS
    packed = ops.stem_pack(w)
    full = x0 if x1 is None else torch.cat([x0, x1], 1)
    for scale, shift in ((1.0, 0.0), (2.0, -1.0)):
        n0 = ops.launch_count()
        y = ops.conv_stem(x0, x1, w, b, in_scale=scale, in_shift=shift, packed=packed)
        one_launch = cin <= 3 and cout in (64, 128)
        assert ops.launch_count() - n0 == (1 if one_launch else 2)   # else: im2col + one GEMM launch
        ref32 = F.conv2d(full * scale + shift, w, b, padding=1)
        ref16 = F.conv2d(_bf16r(full * scale + shift), _bf16r(w), b, padding=1)
        assert y.shape == ref32.shape and y.dtype == torch.bfloat16
        assert _rel_l2(y.float(), ref16) < 3e-3                  # bf16 output rounding
        assert _rel_l2(y.float(), ref32) < 6e-3
        old = ops.conv_stem(x0, x1, w, b, in_scale=scale, in_shift=shift)   # fp32 CUDA-core kernel
        assert _rel_l2(y.float(), old.float()) < 6e-3
    if cout % 32 == 0:
        assert hasattr(y, "_fm_stats")
        gamma, beta = torch.randn(cout, generator=g).to(DEV), torch.randn(cout, generator=g).to(DEV)
        t_fused = ops.group_norm_table([y], 32, 1e-5, gamma, beta, silu=True)
        t_plain = ops.group_norm_table([y.clone(memory_format=torch.preserve_format)], 32, 1e-5, gamma, beta, silu=True)
        assert float((t_fused.ab - t_plain.ab).abs().max()) < 5e-3 * float(t_plain.ab.abs().max())


@pytest.mark.parametrize("case", [
    (2, 5, 256, [128], 128, [3], 1, True, True, True),           # rolling rows
    (1, 6, 200, [64, 128], 256, [3, 3], 1, True, True, False),   # rolling, two N tiles, ragged W
    (3, 14, 14, [128], 128, [3], 1, True, True, True),           # per-tile kernel, several rows per tile
    (2, 16, 16, [1024], 512, [3], 1, True, False, True),
    (2, 32, 32, [256], 256, [1], 1, True, False, True),          # attention output projection shape (1x1 + residual)
    (3, 7, 7, [128], 128, [3], 1, True, False, False),           # several images per tile
], ids=lambda c: "B{}_{}x{}_cout{}".format(c[0], c[1], c[2], c[4]))
def test_conv2d_upsampled_store(case):
    """conv with the nearest-2x of a following UpsampleND folded into its store == interpolate(conv)."""
    B, H, W, cins, cout, ks, stride, bias, addvec, residual = case
    g = torch.Generator(device="cpu").manual_seed(5)
    xs = [_bf16r(torch.randn(B, c, H, W, generator=g)).to(DEV) for c in cins]
    ws = [_bf16r(torch.randn(cout, c, k, k, generator=g) / math.sqrt(c * k * k)).to(DEV) for c, k in zip(cins, ks)]
    bvec = torch.randn(cout, generator=g).to(DEV)
    av = torch.randn(B, cout, generator=g).to(DEV) if addvec else None
    res = _bf16r(torch.randn(B, cout, H, W, generator=g)).to(DEV) if residual else None
    packed = ops.pack_conv_weight([(w, 0, c) for w, c in zip(ws, cins)])
    kw = dict(bias=bvec, addvec=av, residual=_nhwc(res) if residual else None)
    plain = ops.conv2d([_nhwc(x) for x in xs], packed, **kw)
    up = ops.conv2d([_nhwc(x) for x in xs], packed, upsample_out=True, **kw)
    assert up.shape == (B, cout, 2 * H, 2 * W)
    assert torch.equal(up.float(), F.interpolate(plain.float(), scale_factor=2, mode="nearest"))


def test_conv2d_rolling_random_shapes():
    """Seeded random sweep over the rolling-row kernel's argument space (ragged widths, odd heights / batches, 1-3 K
    segments with and without the fused operand transform, residual, per-sample vector, upsampled store, fused
    statistics) against the fp32 reference."""
    import random

    rnd = random.Random(1234)
    g = torch.Generator(device="cpu").manual_seed(99)
    for it in range(24):
        B = rnd.choice([1, 2, 3, 5])
        H = rnd.choice([1, 2, 7, 16, 33])
        W = rnd.choice([65, 100, 128, 130, 200, 256, 300])
        nseg = rnd.choice([1, 1, 2, 3])
        cins = [rnd.choice([64, 128, 192]) for _ in range(nseg)]
        ks = [3] + [rnd.choice([1, 3]) for _ in range(nseg - 1)]
        cout = rnd.choice([64, 128, 192, 256])
        use_norm = rnd.random() < 0.6
        residual = rnd.random() < 0.5
        addvec = rnd.random() < 0.5
        up = rnd.random() < 0.3
        xs = [_bf16r(torch.randn(B, c, H, W, generator=g) * 1.2 + 0.1).to(DEV) for c in cins]
        ws = [_bf16r(torch.randn(cout, c, k, k, generator=g) / math.sqrt(c * k * k)).to(DEV) for c, k in zip(cins, ks)]
        bvec = torch.randn(cout, generator=g).to(DEV)
        av = torch.randn(B, cout, generator=g).to(DEV) if addvec else None
        res = _bf16r(torch.randn(B, cout, H, W, generator=g)).to(DEV) if residual else None
        norm = None
        ys = list(xs)
        if use_norm:
            normed = [k == 3 and rnd.random() < 0.8 for k in ks]
            normed[0] = True
            ctot = sum(c for c, n in zip(cins, normed) if n)
            ab = torch.empty(B, 2, ctot)
            ab[:, 0] = torch.rand(B, ctot, generator=g) + 0.5
            ab[:, 1] = torch.randn(B, ctot, generator=g) * 0.5
            ab = ab.to(DEV)
            table = ops.NormTable(ab, True)
            norm, off = [], 0
            for i, (c, n) in enumerate(zip(cins, normed)):
                if n:
                    y = xs[i] * ab[:, 0, off:off + c, None, None] + ab[:, 1, off:off + c, None, None]
                    ys[i] = _bf16r(F.silu(y))
                    norm.append((table, off))
                    off += c
                else:
                    norm.append(None)
        ref = torch.zeros(B, cout, H, W, device=DEV)
        for y, w, k in zip(ys, ws, ks):
            ref = ref + F.conv2d(y, w, None, padding=k // 2)
        ref = ref + bvec.view(1, -1, 1, 1)
        if addvec:
            ref = ref + av.view(B, cout, 1, 1)
        if residual:
            ref = ref + res
        packed = ops.pack_conv_weight([(w, 0, c) for w, c in zip(ws, cins)])
        out = ops.conv2d([_nhwc(x) for x in xs], packed, bias=bvec, addvec=av, residual=_nhwc(res) if residual else None,
                         norm=norm, upsample_out=up, want_stats=not up)
        if up:
            ref = F.interpolate(ref, scale_factor=2, mode="nearest")
        err = _rel_l2(out.float(), ref)
        tag = (it, B, H, W, cins, ks, cout, use_norm, residual, addvec, up)
        assert err < 6e-3, (tag, err)
        if not up and hasattr(out, "_fm_stats"):
            gamma = torch.ones(cout, device=DEV)
            beta = torch.zeros(cout, device=DEV)
            tab = ops.group_norm_table([out], 32, 1e-5, gamma, beta, silu=False)
            y = out.float() * tab.ab[:, 0, :, None, None] + tab.ab[:, 1, :, None, None]
            want = F.group_norm(out.float(), 32, None, None, 1e-5)
            assert float((y - want).abs().max()) < 2e-2, tag
