"""Training-step kernels against torch fp32 autograd of the same op on the same seeded inputs.

Tolerances: bf16 activations / fp32 accumulation, so gradients are compared by relative L2 (<= 1e-2, the north star's
per-step tolerance for bf16 tensors); the optimiser update and the MSE are fp32 and compared tightly."""
import math

import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def nhwc(x):
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


@pytest.fixture(scope="module")
def F():
    from fmdm_b200.training import functions
    return functions


@pytest.mark.parametrize("cin,cout,k,stride,hw,b", [
    (128, 128, 3, 1, 32, 2), (64, 256, 3, 1, 16, 4), (256, 128, 1, 1, 16, 2), (128, 128, 3, 2, 32, 2),
    (512, 512, 3, 1, 8, 2), (128, 1536, 1, 1, 8, 4), (192, 64, 3, 1, 16, 1),
])
def test_conv_backward(F, cin, cout, k, stride, hw, b):
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(b, cin, hw, hw, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    bias = torch.randn(cout, device=dev)
    addvec = torch.randn(b, cout, device=dev)
    ho = hw // stride
    res = torch.randn(b, cout, ho, ho, device=dev)
    gy = torch.randn(b, cout, ho, ho, device=dev)
    xb = nhwc(x).requires_grad_(True)
    wp = (w if k == 3 else w.reshape(cout, cin)).clone().requires_grad_(True)
    bp, ap = bias.clone().requires_grad_(True), addvec.clone().requires_grad_(True)
    rp = nhwc(res).requires_grad_(True)
    y = F.conv([xb], [(wp, 0, cin)], bias=bp, stride=stride, addvec=ap, residual=rp)
    y.backward(nhwc(gy))
    # reference: fp32 autograd on the bf16-rounded operands
    xr = xb.detach().float().requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    br, ar = bias.clone().requires_grad_(True), addvec.clone().requires_grad_(True)
    rr = rp.detach().float().requires_grad_(True)
    yr = TF.conv2d(xr, wr, br, stride=stride, padding=k // 2) + ar[:, :, None, None] + rr
    yr.backward(nhwc(gy).float())
    assert rel_l2(y, yr) < 1e-2
    assert rel_l2(xb.grad, xr.grad) < 1e-2
    assert rel_l2(wp.grad.reshape(wr.shape), wr.grad) < 1e-2
    assert rel_l2(bp.grad, br.grad) < 1e-2
    assert rel_l2(ap.grad, ar.grad) < 1e-2
    assert rel_l2(rp.grad, rr.grad) < 1e-2


def test_conv_backward_two_segments_shared_and_split_weights(F):
    """conv2 (3x3) + 1x1 skip over a second source accumulate into one output; a virtual concat slices one weight."""
    torch.manual_seed(1)
    dev = "cuda"
    b, hw = 2, 16
    h = nhwc(torch.randn(b, 128, hw, hw, device=dev)).requires_grad_(True)
    x = nhwc(torch.randn(b, 256, hw, hw, device=dev)).requires_grad_(True)
    w2 = (torch.randn(128, 128, 3, 3, device=dev) * 0.05).requires_grad_(True)
    ws = (torch.randn(128, 256, device=dev) * 0.05).requires_grad_(True)
    gy = nhwc(torch.randn(b, 128, hw, hw, device=dev))
    y = F.conv([h, x], [(w2, 0, 128), (ws, 0, 256)])
    y.backward(gy)
    hr, xr = h.detach().float().requires_grad_(True), x.detach().float().requires_grad_(True)
    w2r = w2.detach().to(torch.bfloat16).float().requires_grad_(True)
    wsr = ws.detach().to(torch.bfloat16).float().requires_grad_(True)
    yr = TF.conv2d(hr, w2r, padding=1) + TF.conv2d(xr, wsr[:, :, None, None])
    yr.backward(gy.float())
    for got, ref in [(y, yr), (h.grad, hr.grad), (x.grad, xr.grad), (w2.grad, w2r.grad), (ws.grad, wsr.grad)]:
        assert rel_l2(got, ref) < 1e-2
    # virtual concat: two sources slice ONE weight
    a = nhwc(torch.randn(b, 128, hw, hw, device=dev)).requires_grad_(True)
    c = nhwc(torch.randn(b, 64, hw, hw, device=dev)).requires_grad_(True)
    w = (torch.randn(128, 192, 3, 3, device=dev) * 0.05).requires_grad_(True)
    y = F.conv([a, c], [(w, 0, 128), (w, 128, 64)])
    y.backward(gy)
    ar, cr = a.detach().float().requires_grad_(True), c.detach().float().requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    yr = TF.conv2d(torch.cat([ar, cr], 1), wr, padding=1)
    yr.backward(gy.float())
    for got, ref in [(y, yr), (a.grad, ar.grad), (c.grad, cr.grad), (w.grad, wr.grad)]:
        assert rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("c,groups,hw,b,silu,ss", [
    (128, 32, 32, 2, True, False), (256, 32, 16, 3, True, True), (512, 32, 8, 2, False, False),
    (64, 8, 16, 2, True, True), (384, 32, 16, 2, True, False),
])
def test_group_norm_backward(F, c, groups, hw, b, silu, ss):
    torch.manual_seed(2)
    dev = "cuda"
    x = nhwc(torch.randn(b, c, hw, hw, device=dev) * 1.5 + 0.3).requires_grad_(True)
    gamma = (torch.rand(c, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, device=dev) * 0.2).requires_grad_(True)
    sst = (torch.randn(b, 2 * c, device=dev) * 0.3).requires_grad_(True) if ss else None
    gy = nhwc(torch.randn(b, c, hw, hw, device=dev))
    y = F.group_norm(x, gamma, beta, groups=groups, eps=1e-5, silu=silu, scale_shift=sst)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    sr = sst.detach().clone().requires_grad_(True) if ss else None
    yr = TF.group_norm(xr, groups, gr, br, 1e-5)
    if ss:
        yr = yr * (1 + sr[:, :c, None, None]) + sr[:, c:, None, None]
    if silu:
        yr = TF.silu(yr)
    yr.backward(gy.float())
    assert rel_l2(y, yr) < 1e-2
    assert rel_l2(x.grad, xr.grad) < 1e-2
    assert rel_l2(gamma.grad, gr.grad) < 1e-2
    assert rel_l2(beta.grad, br.grad) < 1e-2
    if ss:
        assert rel_l2(sst.grad, sr.grad) < 1e-2


def test_group_norm_backward_two_sources(F):
    """decoder concat kept virtual: GroupNorm over (hidden, skip) without materialising the concat."""
    torch.manual_seed(21)
    dev = "cuda"
    b, hw, c0, c1 = 2, 16, 256, 128
    x0 = nhwc(torch.randn(b, c0, hw, hw, device=dev) + 0.5).requires_grad_(True)
    x1 = nhwc(torch.randn(b, c1, hw, hw, device=dev) * 2).requires_grad_(True)
    gamma = (torch.rand(c0 + c1, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(c0 + c1, device=dev) * 0.2).requires_grad_(True)
    gy = nhwc(torch.randn(b, c0 + c1, hw, hw, device=dev))
    y = F.group_norm((x0, x1), gamma, beta, groups=32, eps=1e-5, silu=True)
    y.backward(gy)
    r0, r1 = x0.detach().float().requires_grad_(True), x1.detach().float().requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yr = TF.silu(TF.group_norm(torch.cat([r0, r1], 1), 32, gr, br, 1e-5))
    yr.backward(gy.float())
    for got, ref in [(y, yr), (x0.grad, r0.grad), (x1.grad, r1.grad), (gamma.grad, gr.grad), (beta.grad, br.grad)]:
        assert rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("c,heads,hw,b", [(512, 64, 16, 2), (512, 64, 8, 3), (128, 8, 16, 2), (256, 4, 8, 2), (64, 8, 5, 2),
                                          (64, 8, 32, 1)])
def test_attention_backward(F, c, heads, hw, b):
    torch.manual_seed(3)
    dev = "cuda"
    qkv = nhwc(torch.randn(b, 3 * c, hw, hw, device=dev)).requires_grad_(True)
    gy = nhwc(torch.randn(b, c, hw, hw, device=dev))
    y = F.attention_qkv(qkv, heads)
    y.backward(gy)
    t, hd = hw * hw, c // heads
    qr = qkv.detach().float().requires_grad_(True)
    flat = qr.permute(0, 2, 3, 1).reshape(b, t, 3, heads, hd)
    q, k, v = [flat[:, :, i].transpose(1, 2) for i in range(3)]
    o = TF.scaled_dot_product_attention(q, k, v)                      # [b][heads][t][hd]
    yr = o.transpose(1, 2).reshape(b, hw, hw, c).permute(0, 3, 1, 2)
    yr.backward(gy.float())
    assert rel_l2(y, yr) < 1e-2
    assert rel_l2(qkv.grad, qr.grad) < 2e-2


def test_linear_backward(F):
    torch.manual_seed(4)
    dev = "cuda"
    for silu_in in (False, True):
        x = torch.randn(8, 128, device=dev, requires_grad=True)
        w = (torch.randn(512, 128, device=dev) * 0.1).requires_grad_(True)
        bias = torch.randn(512, device=dev, requires_grad=True)
        gy = torch.randn(8, 512, device=dev)
        y = F.linear(x, w, bias, silu_in=silu_in)
        y.backward(gy)
        xr, wr, br = [z.detach().clone().requires_grad_(True) for z in (x, w, bias)]
        yr = TF.linear(TF.silu(xr) if silu_in else xr, wr, br)
        yr.backward(gy)
        for got, ref in [(y, yr), (x.grad, xr.grad), (w.grad, wr.grad), (bias.grad, br.grad)]:
            assert rel_l2(got, ref) < 1e-5


def test_stem_and_head_backward(F):
    torch.manual_seed(5)
    dev = "cuda"
    b, hw = 2, 64
    x0, x1 = torch.randn(b, 1, hw, hw, device=dev), torch.randn(b, 1, hw, hw, device=dev)
    w = (torch.randn(128, 2, 3, 3, device=dev) * 0.2).requires_grad_(True)
    bias = torch.randn(128, device=dev, requires_grad=True)
    gy = nhwc(torch.randn(b, 128, hw, hw, device=dev))
    y = F.conv_stem(x0, x1, w, bias, in_scale=2.0, in_shift=-1.0)
    y.backward(gy)
    wr, br = w.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    yr = TF.conv2d(2 * torch.cat([x0, x1], 1) - 1, wr, br, padding=1)
    yr.backward(gy.float())
    # the weight gradient is the tensor-core GEMM dY^T x im2col(x) with x rounded to bf16, as the conv inputs are under
    # bf16 autocast (measured 1.6e-3); against the bf16-rounded input it is tight
    assert rel_l2(y, yr) < 1e-2 and rel_l2(w.grad, wr.grad) < 4e-3 and rel_l2(bias.grad, br.grad) < 1e-3
    wq = w.detach().clone().requires_grad_(True)
    TF.conv2d((2 * torch.cat([x0, x1], 1) - 1).to(torch.bfloat16).float(), wq, None, padding=1).backward(gy.float())
    assert rel_l2(w.grad, wq.grad) < 2e-4
    # head: bf16 NHWC -> fp32 NCHW, one output channel
    a = nhwc(torch.randn(b, 128, hw, hw, device=dev)).requires_grad_(True)
    wh = (torch.randn(1, 128, 3, 3, device=dev) * 0.1).requires_grad_(True)
    bh = torch.randn(1, device=dev, requires_grad=True)
    g = torch.randn(b, 1, hw, hw, device=dev)
    y = F.conv_head(a, wh, bh)
    y.backward(g)
    ar = a.detach().float().requires_grad_(True)
    whr, bhr = wh.detach().clone().requires_grad_(True), bh.detach().clone().requires_grad_(True)
    yr = TF.conv2d(ar, whr, bhr, padding=1)
    yr.backward(g)
    assert rel_l2(y, yr) < 1e-2
    # dY enters both gradient GEMMs rounded to bf16 (measured: weight gradient 1.6e-3)
    assert rel_l2(a.grad, ar.grad) < 1e-2 and rel_l2(wh.grad, whr.grad) < 4e-3 and rel_l2(bh.grad, bhr.grad) < 1e-4
    whq = wh.detach().clone().requires_grad_(True)
    TF.conv2d(a.detach().float(), whq, None, padding=1).backward(g.to(torch.bfloat16).float())
    assert rel_l2(wh.grad, whq.grad) < 2e-4


def test_upsample_backward_and_mse(F):
    torch.manual_seed(6)
    dev = "cuda"
    x = nhwc(torch.randn(2, 64, 8, 8, device=dev)).requires_grad_(True)
    gy = nhwc(torch.randn(2, 64, 16, 16, device=dev))
    F.upsample_nearest2x(x).backward(gy)
    xr = x.detach().float().requires_grad_(True)
    TF.interpolate(xr, scale_factor=2, mode="nearest").backward(gy.float())
    assert rel_l2(x.grad, xr.grad) < 1e-2
    pred = torch.randn(4, 1, 32, 32, device=dev, requires_grad=True)
    noise, clean = torch.randn_like(pred), torch.randn_like(pred)
    loss = F.mse_loss(pred, noise, clean)
    (loss * 3.0).backward()
    pr = pred.detach().clone().requires_grad_(True)
    lr = TF.mse_loss(pr, noise - clean)
    (lr * 3.0).backward()
    assert abs(loss.item() - lr.item()) < 1e-6 * abs(lr.item()) + 1e-7
    assert rel_l2(pred.grad, pr.grad) < 1e-6


def test_fused_adamw_matches_torch():
    from fmdm_b200.training import FusedAdamW

    torch.manual_seed(7)
    dev = "cuda"
    shapes = [(128, 2, 3, 3), (128,), (512, 128), (7,), (256, 256, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    pr = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = FusedAdamW(ps, lr=1e-3, weight_decay=0.01)
    ref = torch.optim.AdamW(pr, lr=1e-3, weight_decay=0.01)
    for step in range(5):
        opt.zero_grad()
        ref.zero_grad(set_to_none=True)
        for p, r in zip(ps, pr):
            g = torch.randn_like(p)
            p.grad.add_(g)
            r.grad = g.clone()
        opt.step()
        ref.step()
    for p, r in zip(ps, pr):
        assert torch.allclose(p, r, rtol=2e-6, atol=2e-7), (p - r).abs().max().item()


@pytest.mark.parametrize("two_sources", [False, True])
def test_gradient_slots_fold_residual_and_skip_gradients(F, two_sources):
    """A ResBlock-shaped graph: x feeds a GroupNorm (first consumer) and, further down, the residual input of a conv
    (and, with two sources, the 1x1 skip conv of a virtual concat).  With gradient slots the later consumers park their
    contributions and the GroupNorm backward adds them inside its apply pass; the result must equal torch autograd on
    the same graph, and equal the unfused path up to one bf16 rounding."""
    torch.manual_seed(5)
    dev = "cuda"
    b, c, hw = 2, 128, 16
    leaf = nhwc(torch.randn(b, c, hw, hw, device=dev)).requires_grad_(True)
    leaf2 = nhwc(torch.randn(b, c, hw, hw, device=dev)).requires_grad_(True)
    gamma = (torch.rand(2 * c if two_sources else c, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(2 * c if two_sources else c, device=dev) * 0.2).requires_grad_(True)
    w = (torch.randn(c, 2 * c if two_sources else c, 3, 3, device=dev) * 0.03).requires_grad_(True)
    ws = (torch.randn(c, 2 * c, device=dev) * 0.05).requires_grad_(True)
    gy = nhwc(torch.randn(b, c, hw, hw, device=dev))

    def run(fuse):
        old, F.FUSE_GRAD_ACCUMULATION = F.FUSE_GRAD_ACCUMULATION, fuse
        try:
            for t in (leaf, leaf2, gamma, beta, w, ws):
                t.grad = None
            x = F.upsample_nearest2x(leaf)          # activations (non-leaf): eligible for slots
            x2 = F.upsample_nearest2x(leaf2)
            if two_sources:
                h = F.group_norm((x, x2), gamma, beta, groups=32, eps=1e-5, silu=True)
                y = F.conv([h, x, x2], [(w, 0, 2 * c), (ws, 0, c), (ws, c, c)])
            else:
                h = F.group_norm(x, gamma, beta, groups=32, eps=1e-5, silu=True)
                y = F.conv([h], [(w, 0, c)], residual=x)
            if fuse:
                assert getattr(x, "_fm_slot").expected == 1
            y.backward(F.upsample_nearest2x(gy).detach())
            F.assert_slots_drained()
            return y.detach().clone(), [t.grad.detach().clone() for t in (leaf, leaf2 if two_sources else leaf, gamma, w)]
        finally:
            F.FUSE_GRAD_ACCUMULATION = old

    y1, g1 = run(True)
    y0, g0 = run(False)
    assert torch.equal(y1, y0)
    for a, r in zip(g1, g0):
        assert rel_l2(a, r) < 4e-3
    # torch autograd on the same graph (fp32)
    up = lambda t: TF.interpolate(t.detach().float(), scale_factor=2.0, mode="nearest")
    l1, l2 = leaf.detach().float().requires_grad_(True), leaf2.detach().float().requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    wr, wsr = w.detach().clone().requires_grad_(True), ws.detach().clone().requires_grad_(True)
    xr, x2r = TF.interpolate(l1, scale_factor=2.0), TF.interpolate(l2, scale_factor=2.0)
    if two_sources:
        cat = torch.cat([xr, x2r], 1)
        hr = TF.silu(TF.group_norm(cat, 32, gr, br, 1e-5))
        yr = TF.conv2d(hr, wr, padding=1) + TF.conv2d(cat, wsr[:, :, None, None])
    else:
        hr = TF.silu(TF.group_norm(xr, 32, gr, br, 1e-5))
        yr = TF.conv2d(hr, wr, padding=1) + xr
    yr.backward(up(gy))
    assert rel_l2(g1[0], l1.grad) < 1.2e-2 and rel_l2(g1[2], gr.grad) < 1.2e-2 and rel_l2(g1[3], wr.grad) < 1.2e-2
    if two_sources:
        assert rel_l2(g1[1], l2.grad) < 1.2e-2
    # a long-lived leaf consumed by several forward passes never carries a slot (each backward stands alone)
    for _ in range(2):
        hh = F.group_norm(leaf, gamma[:c].detach().requires_grad_(True), beta[:c].detach().requires_grad_(True),
                          groups=32, eps=1e-5, silu=True)
        assert getattr(leaf, "_fm_slot", None) is None
        hh.backward(gy)


@pytest.mark.parametrize("c,heads,hw,tc,b", [(128, 16, 8, 25, 2), (64, 8, 16, 256, 2), (512, 64, 16, 64, 1)])
def test_cross_attention_backward(F, c, heads, hw, tc, b):
    """Cross-attention (Tq != Tk, head_dim 8; `attention.py:232-274` with context_dim): forward and the gradients of q
    and of the K | V buffer against torch SDPA."""
    torch.manual_seed(8)
    dev = "cuda"
    q = nhwc(torch.randn(b, c, hw, hw, device=dev)).requires_grad_(True)
    kv = torch.randn(b, tc, 2 * c, device=dev).to(torch.bfloat16).requires_grad_(True)
    gy = nhwc(torch.randn(b, c, hw, hw, device=dev))
    y = F.cross_attention(q, kv, heads)
    y.backward(gy)
    t, hd = hw * hw, c // heads
    qr = q.detach().float().requires_grad_(True)
    kvr = kv.detach().float().requires_grad_(True)
    qh = qr.permute(0, 2, 3, 1).reshape(b, t, heads, hd).transpose(1, 2)
    kh = kvr[:, :, :c].reshape(b, tc, heads, hd).transpose(1, 2)
    vh = kvr[:, :, c:].reshape(b, tc, heads, hd).transpose(1, 2)
    o = TF.scaled_dot_product_attention(qh, kh, vh)
    yr = o.transpose(1, 2).reshape(b, hw, hw, c).permute(0, 3, 1, 2)
    yr.backward(gy.float())
    assert rel_l2(y, yr) < 1e-2
    assert rel_l2(q.grad, qr.grad) < 2e-2 and rel_l2(kv.grad, kvr.grad) < 2e-2


@pytest.mark.parametrize("cc,groups,tc,o,b", [(4, 4, 64, 256, 2), (8, 8, 50, 128, 3), (16, 16, 256, 1024, 2)])
def test_context_kv_backward(F, cc, groups, tc, o, b):
    """The cross-attention context path (GroupNorm over the context tokens + K | V projection): gradients of the
    projection and of the context GroupNorm's affine against torch autograd (the context itself is data)."""
    torch.manual_seed(9)
    dev = "cuda"
    tokens = torch.randn(b, cc, tc, device=dev) * 1.5 + 0.2
    gamma = (torch.rand(cc, device=dev) + 0.5).requires_grad_(True)
    beta = (torch.randn(cc, device=dev) * 0.3).requires_grad_(True)
    w = (torch.randn(o, cc, device=dev) * 0.3).requires_grad_(True)
    bias = torch.randn(o, device=dev, requires_grad=True)
    gy = torch.randn(b, tc, o, device=dev).to(torch.bfloat16)
    kv = F.context_kv(tokens, gamma, beta, w, bias, groups=groups, eps=1e-5)
    kv.backward(gy)
    gr, br, wr, bir = [z.detach().clone().requires_grad_(True) for z in (gamma, beta, w, bias)]
    n = TF.group_norm(tokens, groups, gr, br, 1e-5)                         # (b, cc, tc)
    ref = torch.einsum("oc,bct->bto", wr, n) + bir
    ref.backward(gy.float())
    assert rel_l2(kv, ref) < 1e-2
    for got, want in ((w.grad, wr.grad), (bias.grad, bir.grad), (gamma.grad, gr.grad), (beta.grad, br.grad)):
        assert rel_l2(got, want) < 1e-4


def _raw_heads(t_cm, heads, dh, parts):
    """(b, parts*inner, T) buffer raw-reshaped to (b, heads, T, parts*dh) and chunked, as `attention.py:111-115` does."""
    b, _, t = t_cm.shape
    return t_cm.reshape(b, heads, t, parts * dh).chunk(parts, dim=-1)


@pytest.mark.parametrize("heads,dh,t,tc,b,linear", [(4, 64, 64, 25, 2, False), (4, 16, 100, 64, 2, False),
                                                    (2, 32, 49, 49, 2, True), (4, 64, 256, 64, 1, True),
                                                    (8, 8, 64, 30, 2, True)])
def test_cross_attention_raw_backward(F, heads, dh, t, tc, b, linear):
    """SpatialCrossAttention's raw-reshape attention (softmax at every head_dim with Tq != Tk, and LinearQKVAttention):
    forward and gradients of the q / kv buffers against torch."""
    torch.manual_seed(10)
    dev = "cuda"
    inner = heads * dh
    q_cm = torch.randn(b, inner, t, device=dev).to(torch.bfloat16).requires_grad_(True)
    kv_cm = torch.randn(b, 2 * inner, tc, device=dev).to(torch.bfloat16).requires_grad_(True)
    gy = torch.randn(b, heads, t, dh, device=dev).to(torch.bfloat16)
    y = F.cross_attention_raw(q_cm, kv_cm, heads, dh, linear=linear)
    y.backward(gy)
    qr, kvr = q_cm.detach().float().requires_grad_(True), kv_cm.detach().float().requires_grad_(True)
    (qh,) = _raw_heads(qr, heads, dh, 1)
    kh, vh = _raw_heads(kvr, heads, dh, 2)
    if linear:
        ks, qs = kh.softmax(dim=-2), qh.softmax(dim=-1)
        ctx = torch.einsum("...nd,...ne->...de", ks, vh) / (ks.sum(dim=-2).unsqueeze(-1) + 1e-6)
        ref = torch.einsum("...nd,...de->...ne", qs, ctx)
    else:
        ref = TF.scaled_dot_product_attention(qh, kh, vh)
    ref.backward(gy.float())
    assert rel_l2(y, ref) < 1e-2
    assert rel_l2(q_cm.grad, qr.grad) < 2e-2 and rel_l2(kv_cm.grad, kvr.grad) < 2e-2


@pytest.mark.parametrize("heads,dh,t,b", [(4, 64, 64, 2), (2, 32, 100, 2), (8, 8, 256, 1)])
def test_linear_self_attention_backward(F, heads, dh, t, b):
    torch.manual_seed(11)
    dev = "cuda"
    inner = heads * dh
    qkv = torch.randn(b, 3 * inner, t, device=dev).to(torch.bfloat16).requires_grad_(True)
    gy = torch.randn(b, heads, t, dh, device=dev).to(torch.bfloat16)
    y = F.attention_raw(qkv, heads, dh, linear=True)
    y.backward(gy)
    r = qkv.detach().float().requires_grad_(True)
    qh, kh, vh = _raw_heads(r, heads, dh, 3)
    ks, qs = kh.softmax(dim=-2), qh.softmax(dim=-1)
    ctx = torch.einsum("...nd,...ne->...de", ks, vh) / (ks.sum(dim=-2).unsqueeze(-1) + 1e-6)
    ref = torch.einsum("...nd,...de->...ne", qs, ctx)
    ref.backward(gy.float())
    assert rel_l2(y, ref) < 1e-2 and rel_l2(qkv.grad, r.grad) < 2e-2
