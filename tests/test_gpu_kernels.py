"""GPU parity tests of every kernel family, called through the C ABI (medicalseg_b200.ops -> libmedseg_b200.so).

Checker = torch reference ops on the same (rounded) inputs; tolerances are stated per test:
  f32 storage : relative max error <= 2e-5 (1e-3 for long f32 reductions whose reference is itself f32)
  bf16 storage: relative max error <= 1.2e-2 (one bf16 rounding of the output is 2^-8 = 3.9e-3 relative)
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

F32_TOL, BF16_TOL = 2e-5, 1.2e-2


def _imp():
    from medicalseg_b200 import ops
    from medicalseg_b200.ops import B8
    return ops, B8


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def tol(dt):
    return F32_TOL if dt == torch.float32 else BF16_TOL


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layout_roundtrip_and_padding(dt):
    ops, B8 = _imp()
    x = torch.randn(2, 20, 5, 6, 7, device="cuda")
    b = B8.from_ncdhw(x, dt, 24)
    ref = x if dt == torch.float32 else x.bfloat16().float()
    assert torch.equal(b.to_ncdhw(20), ref)
    assert float(b.to_ncdhw(24)[:, 20:].abs().max()) == 0.0


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("groups", ["batch", "instance"])
def test_bn_prelu_residual_fwd_bwd(dt, groups):
    ops, B8 = _imp()
    torch.manual_seed(0)
    n, c, dims = 2, 16, (6, 7, 9)
    g = 1 if groups == "batch" else n
    y = torch.randn(n, c, *dims, device="cuda") * 2 + 0.5
    r = torch.randn(n, c, *dims, device="cuda")
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    a1, a2 = torch.rand(c, device="cuda") * 0.5, torch.rand(c, device="cuda") * 0.5
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    yb, rb = B8.from_ncdhw(y, dt), B8.from_ncdhw(r, dt)
    yq, rq = yb.to_ncdhw().requires_grad_(True), rb.to_ncdhw().requires_grad_(True)
    sums = torch.zeros(2 * g * c, dtype=torch.float64, device="cuda")
    ops.bn_stats(yb, g, sums)
    bnbuf = torch.empty(4 * g * c, device="cuda")
    count = yb.s * (n if g == 1 else 1)
    ops.bn_finalize(sums, count, gamma, beta, rm, rv, 0.9, 1e-5, True, c, g, bnbuf)
    out = B8(n, c, dims, dt, device="cuda")
    ops.bn_act_fwd(yb, out, rb, None, 0, bnbuf, a1, a2, g)
    params = [t.clone().requires_grad_(True) for t in (gamma, beta, a1, a2)]
    g_, b_, a1_, a2_ = params
    red_dims = (0, 2, 3, 4) if g == 1 else (2, 3, 4)
    mean = yq.mean(red_dims, keepdim=True)
    var = yq.var(red_dims, unbiased=False, keepdim=True)
    v = lambda p: p.view(1, -1, 1, 1, 1)
    t = (yq - mean) * torch.rsqrt(var + 1e-5) * v(g_) + v(b_)
    act1 = torch.where(t > 0, t, v(a1_) * t)
    t2 = act1 + rq
    ref = torch.where(t2 > 0, t2, v(a2_) * t2)
    assert rel(out.to_ncdhw(), ref.detach()) <= tol(dt)
    # the fused finalize + normalise launch (what VNet uses) gives the same output, bnbuf and running statistics
    rm2, rv2 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    bnbuf2, out2 = torch.zeros(4 * g * c, device="cuda"), B8(n, c, dims, dt, device="cuda")
    ops.bn_fwd_fused(yb, out2, rb, None, 0, sums, count, gamma, beta, rm2, rv2, 0.9, 1e-5, True, bnbuf2, a1, a2, g)
    assert torch.equal(out2.buf, out.buf) and torch.equal(bnbuf2, bnbuf)
    assert torch.equal(rm2, rm) and torch.equal(rv2, rv)
    if g == 1:  # Paddle running-stat convention: 0.9*running + 0.1*batch, biased variance
        assert float((rm - 0.1 * mean.flatten()).abs().max()) < 1e-6
        assert float((rv - (0.9 + 0.1 * var.flatten())).abs().max()) < 1e-5
    go = torch.randn_like(ref)
    gob = B8.from_ncdhw(go, dt)
    ref.backward(gob.to_ncdhw())
    red = torch.zeros(4 * g * c, dtype=torch.float64, device="cuda")
    ops.bn_act_bwd_reduce(yb, rb, None, 0, gob, bnbuf, a1, a2, g, red)
    dy, dres = B8(n, c, dims, dt, device="cuda"), B8(n, c, dims, dt, device="cuda")
    grads = [torch.zeros(c, device="cuda") for _ in range(4)]
    ops.bn_act_bwd_apply(yb, rb, None, 0, gob, bnbuf, a1, a2, red, count, True, dy, dres, False, *grads, g)
    assert rel(dy.to_ncdhw(), yq.grad) <= tol(dt)
    assert rel(dres.to_ncdhw(), rq.grad) <= tol(dt)
    for got, p in zip(grads, params):
        assert rel(got, p.grad) <= max(tol(dt), 1e-5)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_in_tr_tile_add(dt):
    """InputTransition: PReLU(BN(conv(x)) + tile(x,16)) (vnet.py:74-79)"""
    ops, B8 = _imp()
    torch.manual_seed(1)
    n, dims = 2, (9, 11, 37)
    x = torch.rand(n, 1, *dims, device="cuda")
    w = torch.randn(16, 1, 5, 5, 5, device="cuda") * 0.1
    b = torch.randn(16, device="cuda")
    ref = F.conv3d(x, w, b, padding=2)
    out = B8(n, 16, dims, dt, device="cuda")
    sums = torch.zeros(32, dtype=torch.float64, device="cuda")
    ops.conv_in_fwd(x, w, b, out, 1, sums)
    o = out.to_ncdhw()
    assert rel(o, ref) <= tol(dt)
    assert float((sums[:16] - o.double().sum((0, 2, 3, 4))).abs().max()) < 1e-2
    gamma, beta, a1 = torch.ones(16, device="cuda"), torch.zeros(16, device="cuda"), torch.full((16,), 0.25, device="cuda")
    rm, rv = torch.zeros(16, device="cuda"), torch.ones(16, device="cuda")
    bnbuf = torch.empty(64, device="cuda")
    ops.bn_finalize(sums, n * out.s, gamma, beta, rm, rv, 0.9, 1e-5, True, 16, 1, bnbuf)
    act = B8(n, 16, dims, dt, device="cuda")
    ops.bn_act_fwd(out, act, None, x, 1, bnbuf, a1, None, 1)
    bn = F.batch_norm(o, None, None, gamma, beta, True, 0.1, 1e-5)
    t = bn + x.repeat(1, 16, 1, 1, 1)
    refa = torch.where(t > 0, t, 0.25 * t)
    assert rel(act.to_ncdhw(), refa) <= tol(dt)
    dy = torch.randn(n, 16, *dims, device="cuda")
    dyb = B8.from_ncdhw(dy, dt)
    dw, dbias = torch.zeros_like(w), torch.zeros_like(b)
    ops.conv_in_wgrad(x, dyb, dw, dbias)
    dw_ref = torch.nn.grad.conv3d_weight(x, w.shape, dyb.to_ncdhw(), padding=2)
    assert rel(dw, dw_ref) <= 1e-3
    assert rel(dbias, dyb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5


CASES = [((2, 2, 2), (2, 2, 2), 16, 32, (8, 10, 12)),   # isotropic k=s (lung config)
         ((2, 2, 4), (2, 2, 1), 16, 32, (8, 8, 12)),    # MRI level 0: overlap along the last axis
         ((2, 2, 2), (2, 2, 1), 32, 64, (6, 8, 9))]     # MRI level 1


@pytest.mark.parametrize("k,s,ci,co,dims", CASES)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_down_conv_and_up_conv(k, s, ci, co, dims, dt):
    ops, B8 = _imp()
    torch.manual_seed(2)
    n = 2
    x = torch.randn(n, ci, *dims, device="cuda")
    w = torch.randn(co, ci, *k, device="cuda") * 0.1
    b = torch.randn(co, device="cuda")
    xb = B8.from_ncdhw(x, dt)
    xq = xb.to_ncdhw().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    ref = F.conv3d(xq, w_, b, stride=s)
    od = ref.shape[2:]
    out = B8(n, co, od, dt, device="cuda")
    ops.conv_strided_fwd(xb, w, b, out, k, s, (0, 0, 0), 1, None)
    assert rel(out.to_ncdhw(), ref.detach()) <= max(tol(dt), 3e-4)
    dyb = B8.from_ncdhw(torch.randn_like(ref), dt)
    ref.backward(dyb.to_ncdhw())
    dx = B8(n, ci, dims, dt, device="cuda")
    ops.conv_strided_bwd_data(dyb, w, None, dx, k, s, (0, 0, 0), False, 1, None)
    assert rel(dx.to_ncdhw(), xq.grad) <= max(tol(dt), 3e-4)
    dw, dbias = torch.zeros_like(w), torch.zeros_like(b)
    ops.conv_strided_wgrad(xb, dyb, dw, dbias, k, s, (0, 0, 0), False)
    assert rel(dw, w_.grad) <= 1e-3
    assert rel(dbias, dyb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5
    # Conv3DTranspose, weight [Cin_T = co, Cout_T = ci, k]
    wt = torch.randn(co, ci, *k, device="cuda") * 0.1
    bt = torch.randn(ci, device="cuda")
    xsb = B8.from_ncdhw(torch.randn(n, co, *od, device="cuda"), dt)
    xsq = xsb.to_ncdhw().requires_grad_(True)
    wt_ = wt.clone().requires_grad_(True)
    reft = F.conv_transpose3d(xsq, wt_, bt, stride=s)
    outt = B8(n, ci, reft.shape[2:], dt, device="cuda")
    sums = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
    ops.conv_strided_bwd_data(xsb, wt, bt, outt, k, s, (0, 0, 0), False, 1, sums)
    assert rel(outt.to_ncdhw(), reft.detach()) <= max(tol(dt), 3e-4)
    assert float((sums[:ci] - outt.to_ncdhw().double().sum((0, 2, 3, 4))).abs().max()) < 1e-2
    dytb = B8.from_ncdhw(torch.randn_like(reft), dt)
    reft.backward(dytb.to_ncdhw())
    dxs = B8(n, co, od, dt, device="cuda")
    ops.conv_strided_fwd(dytb, wt, None, dxs, k, s, (0, 0, 0), 1, None)
    assert rel(dxs.to_ncdhw(), xsq.grad) <= max(tol(dt), 3e-4)
    dwt, dbt = torch.zeros_like(wt), torch.zeros_like(bt)
    ops.conv_strided_wgrad(dytb, xsb, dwt, dbt, k, s, (0, 0, 0), True)
    assert rel(dwt, wt_.grad) <= 1e-3
    assert rel(dbt, dytb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5


K5_CASES = [(32, 32, (6, 16, 8)), (16, 16, (5, 20, 11)), (64, 64, (4, 16, 16)), (128, 128, (3, 16, 8)),
            (256, 256, (2, 8, 8)), (32, 2, (6, 18, 10)), (32, 32, (9, 7, 13)), (32, 32, (16, 16, 16)),
            (16, 16, (8, 16, 8)), (64, 64, (7, 9, 8)), (32, 16, (19, 16, 8))]


@pytest.mark.parametrize("cin,cout,dims", K5_CASES)
def test_conv_k5_tcgen05_forward(cin, cout, dims):
    """5x5x5 conv on tensor cores vs F.conv3d on the same bf16-rounded operands; ragged tiles included."""
    ops, B8 = _imp()
    torch.manual_seed(3)
    n = 2
    x = torch.randn(n, cin, *dims, device="cuda")
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * (2.0 / (cin * 125)) ** 0.5
    b = torch.randn(cout, device="cuda")
    xb = B8.from_ncdhw(x, torch.bfloat16)
    ref = F.conv3d(xb.to_ncdhw(), w.bfloat16().float(), b, padding=2)
    oc = 16 if cout < 8 else (cout + 7) // 8 * 8
    cout_pad = ops.k5_out_pad(oc)
    packed = torch.empty(ops.k5_packed_bytes(cin, cout_pad), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 0, cin, cout_pad)
    out = B8(n, oc, dims, torch.bfloat16, device="cuda", zero=True)
    sums = torch.zeros(2 * oc, dtype=torch.float64, device="cuda")
    ops.k5_fwd(xb, packed, b, cout, out, False, None, 1, sums)
    o = out.to_ncdhw(cout)
    assert rel(o, ref) <= BF16_TOL
    s_ref = o.double().sum((0, 2, 3, 4))
    assert float((sums[:cout] - s_ref).abs().max()) <= 1e-3 * float(s_ref.abs().max() + 1)
    q_ref = (o.double() ** 2).sum((0, 2, 3, 4))
    assert float((sums[oc:oc + cout] - q_ref).abs().max()) <= 1e-4 * float(q_ref.abs().max() + 1)


def test_conv_k5_tcgen05_dgrad_accumulate_scale():
    ops, B8 = _imp()
    torch.manual_seed(4)
    n, cin, cout, dims = 2, 32, 32, (6, 16, 8)
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.02
    dyb = B8.from_ncdhw(torch.randn(n, cout, *dims, device="cuda"), torch.bfloat16)
    ref = torch.nn.grad.conv3d_input((n, cin, *dims), w.bfloat16().float(), dyb.to_ncdhw(), padding=2)
    cp = ops.k5_out_pad(cin)
    packed = torch.empty(ops.k5_packed_bytes(cout, cp), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 1, cout, cp)
    dx = B8.from_ncdhw(torch.randn(n, cin, *dims, device="cuda"), torch.bfloat16)
    base = dx.to_ncdhw()
    scale = (torch.rand(n, cin, device="cuda") > 0.5).float() * 2
    ops.k5_fwd(dyb, packed, None, cin, dx, True, scale, 1, None)
    assert rel(dx.to_ncdhw(), base + ref * scale.view(n, cin, 1, 1, 1)) <= BF16_TOL


SPLITK_CASES = [(256, 256, (8, 8, 8)), (256, 256, (16, 16, 16)), (128, 128, (16, 16, 16)), (128, 128, (5, 9, 7)),
                (256, 128, (3, 8, 8)), (128, 256, (6, 16, 8))]


@pytest.mark.parametrize("cin,cout,dims", SPLITK_CASES)
def test_conv_k5_split_k_small_volumes(cin, cout, dims):
    """split-K path (msb_conv_k5_fwd_ws) on the small deep-level volumes: forward + BN sums, then the input-gradient
    form with accumulate + channel scale.  The workspace is pure scratch (filled with garbage here) and the result is
    bit-reproducible: every K slice owns a private partial-sum copy that the finalize kernel adds in slice order."""
    ops, B8 = _imp()
    torch.manual_seed(11)
    n = 2
    need = ops.k5_fwd_workspace_bytes(n, cout, dims, cin)
    assert need > 0, "shape is expected to take the split-K path"
    ws = torch.full((need,), 0x7f, dtype=torch.uint8, device="cuda")  # NaN bit patterns: nothing may be read unwritten
    x = torch.randn(n, cin, *dims, device="cuda")
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * (2.0 / (cin * 125)) ** 0.5
    b = torch.randn(cout, device="cuda")
    xb = B8.from_ncdhw(x, torch.bfloat16)
    ref = F.conv3d(xb.to_ncdhw(), w.bfloat16().float(), b, padding=2)
    cout_pad = ops.k5_out_pad(cout)
    packed = torch.empty(ops.k5_packed_bytes(cin, cout_pad), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 0, cin, cout_pad)
    out = B8(n, cout, dims, torch.bfloat16, device="cuda", zero=True)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    ops.k5_fwd(xb, packed, b, cout, out, False, None, 1, sums, ws)
    o = out.to_ncdhw(cout)
    assert rel(o, ref) <= BF16_TOL
    for _ in range(3):  # deterministic: repeated calls give identical bits
        out_r = B8(n, cout, dims, torch.bfloat16, device="cuda", zero=True)
        ops.k5_fwd(xb, packed, b, cout, out_r, False, None, 1, None, ws)
        assert torch.equal(out_r.buf, out.buf)
    s_ref = o.double().sum((0, 2, 3, 4))
    assert float((sums[:cout] - s_ref).abs().max()) <= 1e-3 * float(s_ref.abs().max() + 1)
    q_ref = (o.double() ** 2).sum((0, 2, 3, 4))
    assert float((sums[cout:] - q_ref).abs().max()) <= 1e-4 * float(q_ref.abs().max() + 1)
    # regular path gives the same result up to summation order
    out2 = B8(n, cout, dims, torch.bfloat16, device="cuda", zero=True)
    ops.k5_fwd(xb, packed, b, cout, out2, False, None, 1, None)
    assert rel(out2.to_ncdhw(cout), o) <= BF16_TOL
    # input gradient: dx += scale * conv_T(dy)
    dyb = B8.from_ncdhw(torch.randn(n, cout, *dims, device="cuda"), torch.bfloat16)
    refd = torch.nn.grad.conv3d_input((n, cin, *dims), w.bfloat16().float(), dyb.to_ncdhw(), padding=2)
    cp = ops.k5_out_pad(cin)
    packed_b = torch.empty(ops.k5_packed_bytes(cout, cp), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed_b, cout, cin, 1, cout, cp)
    dx = B8.from_ncdhw(torch.randn(n, cin, *dims, device="cuda"), torch.bfloat16)
    base = dx.to_ncdhw()
    scale = (torch.rand(n, cin, device="cuda") > 0.5).float() * 2
    need_b = ops.k5_fwd_workspace_bytes(n, cin, dims, cout)
    wsb = torch.full((max(need_b, 16),), 0x7f, dtype=torch.uint8, device="cuda")
    ops.k5_fwd(dyb, packed_b, None, cin, dx, True, scale, 1, None, wsb if need_b else None)
    assert rel(dx.to_ncdhw(), base + refd * scale.view(n, cin, 1, 1, 1)) <= BF16_TOL


@pytest.mark.parametrize("cin,cout,dims,res", [(32, 32, (9, 20, 24), True), (64, 64, (5, 16, 16), False),
                                                (16, 32, (8, 16, 16), False), (128, 128, (8, 8, 8), True),
                                                (256, 256, (8, 8, 8), True), (32, 20, (4, 16, 24), False)])
def test_conv_k5_evaluation_epilogue(cin, cout, dims, res):
    """msb_conv_k5_fwd_act (eval-mode LUConv in one kernel): prelu((conv + bias)*scale + shift) and the block tail
    prelu2(. + residual), regular and split-K paths, against an f64 reference on the same bf16-rounded operands."""
    ops, B8 = _imp()
    torch.manual_seed(5)
    n = 2
    x = torch.randn(n, cin, *dims, device="cuda")
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * (2.0 / (cin * 125)) ** 0.5
    b = torch.randn(cout, device="cuda")
    cp = ops.k5_out_pad(cout)
    scale, shift = torch.rand(cp, device="cuda") + 0.5, torch.randn(cp, device="cuda")
    a1, a2 = torch.rand(cp, device="cuda") * 0.5, torch.rand(cp, device="cuda") * 0.5
    xb = B8.from_ncdhw(x, torch.bfloat16)
    rb = B8.from_ncdhw(torch.randn(n, cout, *dims, device="cuda"), torch.bfloat16, c_pad=cp) if res else None
    packed = torch.empty(ops.k5_packed_bytes(cin, cp), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 0, cin, cp)
    bc = lambda v: v[:cout].double().view(1, cout, 1, 1, 1)
    t = F.conv3d(xb.to_ncdhw().double(), w.bfloat16().double(), b.double(), padding=2) * bc(scale) + bc(shift)
    t = torch.where(t > 0, t, t * bc(a1))
    if res:
        t = t + rb.to_ncdhw(cout).double()
        t = torch.where(t > 0, t, t * bc(a2))
    need = ops.k5_fwd_workspace_bytes(n, cp, dims, cin)
    assert (need > 0) == (cin >= 128), "the 8^3 cases are expected to take the split-K path"
    ws = torch.zeros(need, dtype=torch.uint8, device="cuda") if need else None
    out = B8(n, cp, dims, torch.bfloat16, device="cuda", zero=True)
    ops.k5_fwd_act(xb, packed, b, cout, out, scale, shift, a1, rb, a2 if res else None, ws)
    assert rel(out.to_ncdhw(cout).double(), t) <= BF16_TOL
    if cp > cout:  # padded output channels: zero weights and bias -> prelu(shift) (the model pads shift with 0)
        pad_ref = torch.where(shift[cout:] > 0, shift[cout:], shift[cout:] * a1[cout:]).bfloat16().float()
        assert torch.equal(out.to_ncdhw(cp)[:, cout:], pad_ref.view(1, -1, 1, 1, 1).expand(n, -1, *dims))
    with pytest.raises(Exception, match="alpha2"):
        ops.k5_fwd_act(xb, packed, b, cout, out, scale, shift, a1, out, None, ws)


WG_CASES = [(32, 32, (6, 16, 16)), (64, 64, (4, 9, 20)), (128, 128, (3, 8, 16)), (256, 256, (2, 8, 8)),
            (32, 2, (5, 10, 18)), (16, 16, (5, 8, 16)), (32, 32, (3, 13, 9)),
            # enough tiles for the clustered (TMA-multicast) launch: 2-CTA clusters (32 ch), 6-CTA clusters (64 ch)
            (32, 32, (16, 64, 64)), (64, 64, (16, 32, 32)), (32, 32, (9, 70, 50)), (128, 128, (16, 32, 32)),
            (256, 256, (8, 32, 32))]


@pytest.mark.parametrize("cin,cout,dims", WG_CASES)
@pytest.mark.parametrize("clustered", [False, True, "norep"])
def test_conv_k5_tcgen05_wgrad(cin, cout, dims, clustered):
    """clustered: True forces the TMA-multicast cluster launch of the kh-stacked kernel for every shape family (default:
    only the balanced 2-CTA clusters of the 32-channel layers, see conv_k5_wgrad2.cu), False disables it; only the
    shapes with enough tiles take it.  The 32-channel clusters load the leftover kd plane as 4 kw-shifted replicas
    (round 2); "norep" keeps the round-1 form of that group (one valid plane per M block) verified."""
    ops, B8 = _imp()
    from medicalseg_b200 import _lib
    if clustered and dims[0] * dims[1] * dims[2] < 16 * 32 * 32:
        pytest.skip("too few tiles for the cluster launch")
    if clustered == "norep" and cin != 32:
        pytest.skip("the replica form only exists for 32 input channels")
    torch.manual_seed(5)
    n = 2
    dyc = 16 if cout < 8 else (cout + 7) // 8 * 8
    dy = torch.zeros(n, dyc, *dims, device="cuda")
    dy[:, :cout] = torch.randn(n, cout, *dims, device="cuda")
    xb = B8.from_ncdhw(torch.randn(n, cin, *dims, device="cuda"), torch.bfloat16)
    dyb = B8.from_ncdhw(dy, torch.bfloat16)
    # f64 reference (torch's f32 convolution gradients may run in TF32)
    ref = torch.nn.grad.conv3d_weight(xb.to_ncdhw().double(), (cout, cin, 5, 5, 5), dyb.to_ncdhw(cout).double(),
                                      padding=2).float()
    dw = torch.zeros(cout, cin, 5, 5, 5, device="cuda")
    db = torch.zeros(cout, device="cuda")
    ws = torch.empty(ops.k5_wgrad_workspace_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    # 8: clustered per-tap kernel (>= 128 channels); 16: no kw-replicated leftover plane
    _lib.call("msb_debug_set", 6, ((4 | 8 | 16) if clustered == "norep" else (4 | 8)) if clustered else 2)
    try:
        ops.k5_wgrad(xb, dyb, dw, db, cout, cin, ws)
        assert rel(dw, ref) <= 1e-4  # bf16 x bf16 products are exact in f32; only the summation order differs
        assert rel(db, dyb.to_ncdhw(cout).sum((0, 2, 3, 4))) <= 1e-5
        ops.k5_wgrad(xb, dyb, dw, db, cout, cin, ws)  # accumulates (+=)
        assert rel(dw, 2 * ref) <= 1e-4
    finally:
        _lib.call("msb_debug_set", 6, 0)


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (64, 64), (128, 128), (256, 256), (32, 20), (2, 16), (16, 2),
                                      (48, 24), (256, 128), (64, 256)])
@pytest.mark.parametrize("lo_part", [0, 1])
def test_conv_k5_pack_tm_pair_is_bit_identical_to_the_two_single_image_packs(cin, cout, lo_part):
    """msb_conv_k5_pack_tm_pair (one read of the master weight, both operand images - what the train step calls after
    every optimizer step) writes exactly the bytes of msb_conv_k5_pack_tm mode 0 and mode 1, padding included."""
    ops, _ = _imp()
    torch.manual_seed(7)
    w_tm = torch.randn(125, cout, cin, device="cuda")
    f_cin_pad, f_cout_pad = (cin + 15) // 16 * 16, ops.k5_out_pad((cout + 7) // 8 * 8)
    b_cin_pad, b_cout_pad = (cout + 15) // 16 * 16, ops.k5_out_pad((cin + 7) // 8 * 8)
    ref_f = torch.zeros(ops.k5_packed_bytes(f_cin_pad, f_cout_pad), dtype=torch.uint8, device="cuda")
    ref_b = torch.zeros(ops.k5_packed_bytes(b_cin_pad, b_cout_pad), dtype=torch.uint8, device="cuda")
    ops.k5_pack_tm(w_tm, ref_f, cout, cin, 0 | (2 * lo_part), f_cin_pad, f_cout_pad)
    ops.k5_pack_tm(w_tm, ref_b, cout, cin, 1 | (2 * lo_part), b_cin_pad, b_cout_pad)
    got_f, got_b = torch.full_like(ref_f, 0xAB), torch.full_like(ref_b, 0xAB)  # poisoned: every byte must be written
    ops.k5_pack_tm_pair(w_tm, got_f, got_b, cout, cin, lo_part, f_cin_pad, f_cout_pad, b_cin_pad, b_cout_pad)
    assert torch.equal(got_f, ref_f)
    assert torch.equal(got_b, ref_b)


@pytest.mark.parametrize("cin,cout,dims", [(32, 32, (6, 16, 16)), (64, 64, (4, 9, 20)), (128, 128, (3, 8, 16)),
                                           (32, 20, (5, 10, 18))])
def test_conv_k5_tap_major_pack_and_wgrad(cin, cout, dims):
    """tap-major master layout [125][co][ci]: msb_conv_k5_pack_tm builds byte-identical operand images, and
    msb_conv_k5_wgrad_tm accumulates straight into the tap-major gradient (no workspace / unpack) with the same values
    as the Paddle-layout entry point."""
    ops, B8 = _imp()
    torch.manual_seed(6)
    n = 2
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda")
    w_tm = w.reshape(cout, cin, 125).permute(2, 0, 1).contiguous()
    dyc = (cout + 7) // 8 * 8
    for mode, (cin_pad, cout_pad) in ((0, (cin, ops.k5_out_pad(dyc))), (1, ((cout + 15) // 16 * 16, ops.k5_out_pad(cin)))):
        a = torch.zeros(ops.k5_packed_bytes(cin_pad, cout_pad), dtype=torch.uint8, device="cuda")
        b = torch.ones_like(a)
        ops.k5_pack(w, a, cout, cin, mode, cin_pad, cout_pad)
        ops.k5_pack_tm(w_tm, b, cout, cin, mode, cin_pad, cout_pad)
        assert torch.equal(a, b)
    dy = torch.zeros(n, dyc, *dims, device="cuda")
    dy[:, :cout] = torch.randn(n, cout, *dims, device="cuda")
    xb = B8.from_ncdhw(torch.randn(n, cin, *dims, device="cuda"), torch.bfloat16)
    dyb = B8.from_ncdhw(dy, torch.bfloat16)
    dw, db = torch.zeros(cout, cin, 5, 5, 5, device="cuda"), torch.zeros(cout, device="cuda")
    ws = torch.empty(ops.k5_wgrad_workspace_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    ops.k5_wgrad(xb, dyb, dw, db, cout, cin, ws)
    dw_tm, db2 = torch.zeros(125, cout, cin, device="cuda"), torch.zeros(cout, device="cuda")
    ops.k5_wgrad_tm(xb, dyb, dw_tm, db2, cout, cin)
    back = dw_tm.permute(1, 2, 0).reshape(cout, cin, 5, 5, 5)
    assert rel(back, dw) <= 1e-5 and rel(db2, db) <= 1e-6
    ops.k5_wgrad_tm(xb, dyb, dw_tm, None, cout, cin)  # += semantics
    assert rel(dw_tm.permute(1, 2, 0).reshape(cout, cin, 5, 5, 5), 2 * dw) <= 1e-5


@pytest.mark.parametrize("cin,cout,dims", [(32, 32, (9, 16, 16)), (64, 64, (4, 16, 8)), (128, 128, (3, 16, 8))])
def test_conv_k5_three_pass_fp32(cin, cout, dims):
    """3 x bf16 fp32 path at the kernel level: split_hi_lo + lo-part weight images + f32-accumulating epilogue reproduce
    an f64 convolution to ~2^-16 (tolerance 1e-4 of the output range; a single bf16 pass is at 4e-3)."""
    ops, B8 = _imp()
    torch.manual_seed(21)
    n = 2
    x = torch.randn(n, cin, *dims, device="cuda")
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * (2.0 / (cin * 125)) ** 0.5
    b = torch.randn(cout, device="cuda")
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=2).float()
    xb = B8.from_ncdhw(x, torch.float32)
    hi, lo = B8(n, cin, dims, torch.bfloat16, device="cuda"), B8(n, cin, dims, torch.bfloat16, device="cuda")
    ops.split_hi_lo(xb, hi, lo)
    assert rel(hi.to_ncdhw() + lo.to_ncdhw(), x) <= 2e-5
    w_tm = w.reshape(cout, cin, 125).permute(2, 0, 1).contiguous()
    cp = ops.k5_out_pad(cout)
    p_hi = torch.empty(ops.k5_packed_bytes(cin, cp), dtype=torch.uint8, device="cuda")
    p_lo = torch.empty_like(p_hi)
    ops.k5_pack_tm(w_tm, p_hi, cout, cin, 0, cin, cp)
    ops.k5_pack_tm(w_tm, p_lo, cout, cin, 0 | 2, cin, cp)
    out = B8(n, cout, dims, torch.float32, device="cuda", zero=True)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    ops.k5_fwd(hi, p_hi, b, cout, out, False, None, 1, None)
    one_pass = rel(out.to_ncdhw(), ref)
    ops.k5_fwd(lo, p_hi, None, cout, out, True, None, 1, None)
    ops.k5_fwd(hi, p_lo, None, cout, out, True, None, 1, sums)
    o = out.to_ncdhw()
    assert rel(o, ref) <= 1e-4 < one_pass
    s_ref = o.double().sum((0, 2, 3, 4))
    assert float((sums[:cout] - s_ref).abs().max()) <= 1e-3 * float(s_ref.abs().max() + 1)


@pytest.mark.parametrize("ci,co,dims", [(16, 32, (8, 12, 16)), (64, 128, (4, 8, 32)), (128, 256, (4, 4, 16))])
def test_k2s2_tensor_core_wgrad(ci, co, dims):
    """2x2x2 stride-2 weight gradients (space-to-depth + pointwise tcgen05 GEMM) for conv and transposed conv"""
    ops, B8 = _imp()
    torch.manual_seed(9)
    n = 2
    od = tuple(d // 2 for d in dims)
    xb = B8.from_ncdhw(torch.randn(n, ci, *dims, device="cuda"), torch.bfloat16)
    dyb = B8.from_ncdhw(torch.randn(n, co, *od, device="cuda"), torch.bfloat16)
    ws = torch.empty(ops.k2s2_wgrad_workspace_bytes(n, ci, co, dims), dtype=torch.uint8, device="cuda")
    # nn.Conv3D: big = x, small = dy, dw [co][ci][2,2,2]
    ref = torch.nn.grad.conv3d_weight(xb.to_ncdhw(), (co, ci, 2, 2, 2), dyb.to_ncdhw(), stride=2)
    dw, db = torch.zeros(co, ci, 2, 2, 2, device="cuda"), torch.zeros(co, device="cuda")
    ops.k2s2_wgrad(xb, dyb, dw, db, False, ws)
    assert rel(dw, ref) <= 1e-4
    assert rel(db, dyb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5
    # nn.Conv3DTranspose (weight [Cin_T = co, Cout_T = ci, k]): big = dy_out (ci channels), small = x_in (co channels)
    x_in = dyb.to_ncdhw().requires_grad_(False)
    wt = torch.randn(co, ci, 2, 2, 2, device="cuda", requires_grad=True)
    out = F.conv_transpose3d(x_in, wt, stride=2)
    out.backward(xb.to_ncdhw())
    dwt, dbt = torch.zeros(co, ci, 2, 2, 2, device="cuda"), torch.zeros(ci, device="cuda")
    ops.k2s2_wgrad(xb, dyb, dwt, dbt, True, ws)
    assert rel(dwt, wt.grad) <= 1e-4
    assert rel(dbt, xb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5


TC_CASES = [((2, 2, 2), (2, 2, 2), 16, 32, (8, 32, 16)), ((2, 2, 4), (2, 2, 1), 16, 32, (8, 32, 12)),
            ((2, 2, 2), (2, 2, 1), 32, 64, (8, 16, 9)), ((2, 2, 4), (2, 2, 1), 16, 32, (6, 34, 12)),
            ((2, 2, 2), (2, 2, 1), 64, 128, (4, 36, 9)), ((2, 2, 2), (2, 2, 2), 128, 256, (4, 8, 16))]


@pytest.mark.parametrize("k,s,ci,co,dims", TC_CASES)
def test_strided_tensor_core_convs_any_kernel_stride(k, s, ci, co, dims):
    """msb_conv_tc_*: down_conv (gather), up_conv (scatter, incl. overlapping windows along w via the gather-w form),
    their input gradients and weight gradients on tensor cores, vs torch on the same bf16-rounded operands.
    Shapes: the default 2x2x2/stride 2 and the MRISpineSeg kernels (2,2,4)/(2,2,1), (2,2,2)/(2,2,1)."""
    ops, B8 = _imp()
    torch.manual_seed(13)
    n = 2
    od = tuple((d - kk) // ss + 1 for d, kk, ss in zip(dims, k, s))
    taps = k[0] * k[1] * k[2]
    # ---- Conv3D ci -> co (weight [co][ci][k]): forward = gather, input gradient = scatter (accumulating)
    w = torch.randn(co, ci, *k, device="cuda") * (1.0 / (ci * taps)) ** 0.5
    b = torch.randn(co, device="cuda")
    xb = B8.from_ncdhw(torch.randn(n, ci, *dims, device="cuda"), torch.bfloat16)
    xq = xb.to_ncdhw().requires_grad_(True)
    wq = w.bfloat16().float().requires_grad_(True)
    ref = F.conv3d(xq, wq, b, stride=s)
    assert tuple(ref.shape[2:]) == od
    pk = torch.empty(ops.tc_packed_bytes(ci, co, k, s, 0), dtype=torch.uint8, device="cuda")
    ops.tc_pack(w, pk, ci, co, 0, ci, co, k, s)
    out = B8(n, co, od, torch.bfloat16, device="cuda")
    sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    ops.tc_gather(xb, pk, b, co, out, k, s, 1, sums)
    o = out.to_ncdhw()
    assert rel(o, ref.detach()) <= BF16_TOL
    s_ref = o.double().sum((0, 2, 3, 4))
    assert float((sums[:co] - s_ref).abs().max()) <= 1e-3 * float(s_ref.abs().max() + 1)
    dyb = B8.from_ncdhw(torch.randn_like(ref), torch.bfloat16)
    ref.backward(dyb.to_ncdhw())
    pk1 = torch.empty(ops.tc_packed_bytes(co, ci, k, s, 1), dtype=torch.uint8, device="cuda")
    assert pk1.numel() > 0
    ops.tc_pack(w, pk1, co, ci, 1, co, ci, k, s)
    dx = B8.from_ncdhw(torch.randn(n, ci, *dims, device="cuda"), torch.bfloat16)
    base = dx.to_ncdhw()
    ops.tc_scatter(dyb, pk1, None, ci, dx, k, s, True)
    assert rel(dx.to_ncdhw(), base + xq.grad) <= BF16_TOL
    dw, db = torch.zeros_like(w), torch.zeros(co, device="cuda")
    ws = torch.empty(ops.tc_wgrad_workspace_bytes(ci, co, k), dtype=torch.uint8, device="cuda")
    ops.tc_wgrad(xb, dyb, dw, db, k, s, False, ws)
    assert rel(dw, wq.grad) <= 1e-4
    assert rel(db, dyb.to_ncdhw().sum((0, 2, 3, 4))) <= 1e-5
    # ---- Conv3DTranspose co -> ci (weight [co][ci][k]): forward = scatter with bias + BN sums, input gradient = gather
    wt = torch.randn(co, ci, *k, device="cuda") * (1.0 / (co * taps)) ** 0.5
    bt = torch.randn(ci, device="cuda")
    sb = B8.from_ncdhw(torch.randn(n, co, *od, device="cuda"), torch.bfloat16)
    sq_ = sb.to_ncdhw().requires_grad_(True)
    wtq = wt.bfloat16().float().requires_grad_(True)
    reft = F.conv_transpose3d(sq_, wtq, bt, stride=s)
    assert tuple(reft.shape[2:]) == tuple(dims)
    pkt = torch.empty(ops.tc_packed_bytes(co, ci, k, s, 1), dtype=torch.uint8, device="cuda")
    ops.tc_pack(wt, pkt, co, ci, 1, co, ci, k, s)
    up = B8(n, ci, dims, torch.bfloat16, device="cuda", zero=True)
    sums2 = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
    ops.tc_scatter(sb, pkt, bt, ci, up, k, s, False, 1, sums2)
    u = up.to_ncdhw()
    assert rel(u, reft.detach()) <= BF16_TOL
    q_ref = (u.double() ** 2).sum((0, 2, 3, 4))
    assert float((sums2[ci:] - q_ref).abs().max()) <= 1e-4 * float(q_ref.abs().max() + 1)
    dub = B8.from_ncdhw(torch.randn_like(reft), torch.bfloat16)
    reft.backward(dub.to_ncdhw())
    pkg = torch.empty(ops.tc_packed_bytes(ci, co, k, s, 0), dtype=torch.uint8, device="cuda")
    ops.tc_pack(wt, pkg, ci, co, 0, ci, co, k, s)
    dsm = B8(n, co, od, torch.bfloat16, device="cuda")
    ops.tc_gather(dub, pkg, None, co, dsm, k, s)
    assert rel(dsm.to_ncdhw(), sq_.grad) <= BF16_TOL
    dwt = torch.zeros_like(wt)
    ops.tc_wgrad(dub, sb, dwt, None, k, s, True, ws)
    assert rel(dwt, wtq.grad) <= 1e-4


@pytest.mark.parametrize("c", [2, 3, 20])
def test_fused_dice_ce_loss_matches_oracle(c):
    from oracle import vnet_oracle as vo
    from medicalseg_b200.models import losses as L
    torch.manual_seed(6)
    logits = torch.randn(2, c, 8, 9, 10) * 2
    labels = torch.randint(0, c, (2, 8, 9, 10), dtype=torch.int32)
    labels[0, 0, 0, :3] = 255  # ignore_index voxels (CE only)
    lo = logits.clone().requires_grad_(True)
    ll, dice = vo.loss_computation([lo], labels, vo.default_losses())
    sum(ll).backward()
    lg = logits.cuda().requires_grad_(True)
    ours = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    l2, d2 = L.loss_computation([lg], labels.cuda(), ours)
    sum(l2).backward()
    for a, b in zip(ll, l2):
        assert abs(float(a) - float(b)) <= 1e-5  # CE/Dice abs <= 1e-5 in f32
    assert float(np.abs(dice - d2).max()) <= 1e-6
    assert rel(lg.grad.cpu(), lo.grad) <= 1e-5


@pytest.mark.parametrize("c,cpad,dims", [(3, 16, (5, 6, 7)), (2, 8, (9, 17, 20)), (20, 32, (6, 30, 12)), (32, 32, (3, 40, 9))])
def test_conv1x1_head(c, cpad, dims):
    ops, B8 = _imp()
    torch.manual_seed(7)
    n = 2
    a = torch.zeros(n, cpad, *dims, device="cuda")
    a[:, :c] = torch.randn(n, c, *dims, device="cuda")
    w, b = torch.randn(c, c, device="cuda"), torch.randn(c, device="cuda")
    ab = B8.from_ncdhw(a, torch.float32)
    logits = torch.empty(n, c, *dims, device="cuda")
    ops.conv1x1_fwd(ab, w, b, logits, c, c)
    a_ = a[:, :c].double().requires_grad_(True)   # f64 reference (torch's f32 conv may use TF32)
    w_ = w.double().requires_grad_(True)
    b_ = b.double().requires_grad_(True)
    ref = F.conv3d(a_, w_.view(c, c, 1, 1, 1), b_)
    assert rel(logits, ref.detach().float()) <= 1e-5
    dl = torch.randn_like(logits)
    ref.backward(dl.double())
    da = B8(n, cpad, dims, torch.float32, device="cuda")
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.conv1x1_bwd(ab, w, dl, da, dw, db, c, c)
    assert rel(da.to_ncdhw(c), a_.grad.float()) <= 1e-5
    assert rel(dw, w_.grad.float()) <= 1e-4 and rel(db, b_.grad.float()) <= 1e-5
    if cpad > c:
        assert float(da.to_ncdhw(cpad)[:, c:].abs().max()) == 0.0


def test_momentum_step_matches_oracle():
    from medicalseg_b200 import ops
    torch.manual_seed(8)
    p, g, v = torch.randn(1003), torch.randn(1003), torch.randn(1003)
    pc, gc, vc = p.cuda(), g.cuda(), v.cuda()
    ops.momentum_step(pc, gc, vc, 0.01, 0.9, 1e-4, 0.5)
    v_ref = 0.9 * v + (0.5 * g + 1e-4 * p)
    p_ref = p - 0.01 * v_ref
    assert rel(vc.cpu(), v_ref) <= 1e-6 and rel(pc.cpu(), p_ref) <= 1e-6


def test_preprocess_matches_reference_golden():
    """HUnorm / normalize / resample / label_remap vs outputs of the reference's own functions (tests/golden)."""
    from medicalseg_b200 import preprocess as P
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_ref.npz"))
    for name, shp in (("iso", (16, 16, 16)), ("aniso", (16, 24, 12)), ("up", (20, 17, 23)), ("one", (1, 4, 8))):
        assert np.array_equal(P.HUnorm(G[name + "_in"]), G[name + "_hunorm"])  # bit-exact
        r1, sp = P.resample(G[name + "_hunorm"], spacing=(1.0, 0.7, 0.7), new_shape=list(shp), order=1)
        assert r1.dtype == np.float32 and r1.shape == G[name + "_resample1"].shape
        # order 1: abs <= 1e-3 on the 0..255 scale (f32 blend vs SciPy's f64), SURVEY §8d
        assert float(np.abs(r1 - G[name + "_resample1"]).max()) <= 1e-3
        np.testing.assert_allclose(np.asarray(sp), G[name + "_spacing"])
        rf, _ = P.resample(G[name + "_in"], new_shape=list(shp), order=1, pre_op=("hunorm", -1200, 600, -2000))
        assert float(np.abs(rf - G[name + "_resample1"]).max()) <= 1e-3
        r0, _ = P.resample(G[name + "_label"], new_shape=list(shp), order=0)
        assert np.array_equal(r0, G[name + "_resample0"])  # order 0: exact
        vol = np.nan_to_num(G[name + "_in"], nan=0.0)
        assert float(np.abs(P.normalize(vol) - G[name + "_norm_minmax"]).max()) <= 1e-7
        assert float(np.abs(P.normalize(vol, 0, 2650) - G[name + "_norm_fixed"]).max()) <= 1e-7
    assert np.array_equal(P.label_remap(G["remap_in"], {1: 2, 2: 3, 5: 0}), G["remap_out"])
    r, _ = P.resample(G["spacing_in"], spacing=(2.0, 1.5, 0.5), new_spacing=[1.0, 1.0, 1.0], order=1)
    assert r.shape == G["spacing_out"].shape and float(np.abs(r - G["spacing_out"]).max()) <= 1e-3


def test_resample_full_size_properties():
    """size-independent properties at a BASELINE-like size: identity zoom is exact, constant volumes stay constant,
    order-0 output only contains input labels, monotone ramps stay monotone."""
    from medicalseg_b200 import preprocess as P
    x = torch.rand(96, 160, 160, device="cuda") * 255
    same, _ = P.resample(x, new_shape=[96, 160, 160], order=1)
    assert torch.equal(same, x)
    const = torch.full((64, 64, 64), 7.25, device="cuda")
    out, _ = P.resample(const, new_shape=[24, 40, 56], order=1)
    assert float((out - 7.25).abs().max()) == 0.0
    lab = torch.randint(0, 3, (96, 128, 128), device="cuda", dtype=torch.int32)
    o0, _ = P.resample(lab, new_shape=[32, 32, 32], order=0)
    assert set(o0.unique().tolist()) <= {0, 1, 2}
    ramp = torch.arange(256, device="cuda", dtype=torch.float32).view(1, 1, 256).expand(8, 8, 256).contiguous()
    r, _ = P.resample(ramp, new_shape=[8, 8, 64], order=1)
    assert bool((r[..., 1:] > r[..., :-1]).all())


@pytest.mark.parametrize("small,big", [((16, 16, 16), (32, 32, 32)), ((4, 4, 4), (32, 32, 32)), ((8, 8, 4), (64, 64, 12)),
                                       ((5, 7, 3), (12, 20, 9)), ((9, 6, 10), (9, 6, 10)), ((12, 10, 8), (5, 4, 3)),
                                       ((2, 3, 2), (48, 40, 64))])  # last: footprints beyond the cached-weight length
def test_trilinear_resize_and_adjoint(small, big):
    """msb_trilinear_fwd / _bwd (VNetDeepSup heads, vnet_deepsup.py:259-272) vs torch F.interpolate(trilinear,
    align_corners=False) and its autograd adjoint; f32, 2e-5 relative."""
    ops, _ = _imp()
    torch.manual_seed(2)
    x = torch.randn(2, 3, *small, device="cuda", requires_grad=True)
    ref = F.interpolate(x, size=big, mode="trilinear", align_corners=False)
    out = torch.empty(2, 3, *big, device="cuda")
    ops.trilinear_fwd(x.detach(), out)
    assert rel(out, ref.detach()) <= F32_TOL
    g = torch.randn_like(ref)
    ref.backward(g)
    dx = torch.full_like(x, float("nan"))
    ops.trilinear_bwd(g, dx)
    assert rel(dx, x.grad) <= F32_TOL
    # adjoint identity <A x, g> == <x, A^T g> in f64
    lhs = float((out.double() * g.double()).sum())
    rhs = float((x.detach().double() * dx.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_dynamic_tile_scheduler_matches_static_assignment():
    """msb_set_tile_scheduler(1): the persistent kernels fetch tiles from a self-resetting atomic counter (used by the
    data-parallel step, where NCCL's CTAs displace some conv CTAs).  Forward / dgrad and the strided convs must be
    bit-identical to the static round-robin, the weight gradients equal up to the order of their f32 atomics; every
    call is repeated to prove that the counters reset themselves."""
    ops, B8 = _imp()
    from medicalseg_b200 import _lib
    if _lib.load().msb_set_tile_scheduler(1) != 0:
        # the default build compiles the scheduler out (umma.cuh, MSB_DYNAMIC_TILES): the entry point must refuse, loudly
        with pytest.raises(RuntimeError, match="compiled out"):
            _lib.call("msb_set_tile_scheduler", 1)
        _lib.call("msb_set_tile_scheduler", 0)
        pytest.skip("dynamic tile scheduler compiled out (MSB_DYNAMIC_TILES=1 python -m medicalseg_b200.build --force)")
    torch.manual_seed(31)
    n = 2

    def fwd_case(c, dims):
        x = B8.from_ncdhw(torch.randn(n, c, *dims, device="cuda"), torch.bfloat16)
        w = torch.randn(c, c, 5, 5, 5, device="cuda") * 0.02
        cp = ops.k5_out_pad(c)
        packed = torch.empty(ops.k5_packed_bytes(c, cp), dtype=torch.uint8, device="cuda")
        ops.k5_pack(w, packed, c, c, 0, c, cp)
        bias = torch.randn(c, device="cuda")

        def run():
            out = B8(n, c, dims, torch.bfloat16, device="cuda", zero=True)
            sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
            ops.k5_fwd(x, packed, bias, c, out, False, None, 1, sums)
            return out.buf.clone(), sums
        return run

    def wgrad_case(c, dims):
        x = B8.from_ncdhw(torch.randn(n, c, *dims, device="cuda"), torch.bfloat16)
        dy = B8.from_ncdhw(torch.randn(n, c, *dims, device="cuda"), torch.bfloat16)

        def run():
            dw = torch.zeros(125 * c * c, device="cuda")
            ops.k5_wgrad_tm(x, dy, dw, None, c, c)
            return dw, None
        return run

    def k2_case(scatter, accumulate=False):
        a, b, big, small = 32, 16, (24, 32, 32), (12, 16, 16)
        w = torch.randn(a, b, 2, 2, 2, device="cuda") * 0.1
        if scatter:
            x = B8.from_ncdhw(torch.randn(n, a, *small, device="cuda"), torch.bfloat16)
            packed = torch.empty(ops.k2s2_packed_bytes(a, 16), dtype=torch.uint8, device="cuda")
            ops.k2s2_pack(w, packed, a, b, 1, a, 16)
            base = torch.randn(n, b, *big, device="cuda")

            def run():
                out = B8.from_ncdhw(base, torch.bfloat16)
                ops.k2s2_scatter(x, packed, None, b, out, accumulate, 1, None)
                return out.buf.clone(), None
        else:
            x = B8.from_ncdhw(torch.randn(n, b, *big, device="cuda"), torch.bfloat16)
            packed = torch.empty(ops.k2s2_packed_bytes(b, 32), dtype=torch.uint8, device="cuda")
            ops.k2s2_pack(w, packed, b, a, 0, b, 32)

            def run():
                out = B8(n, a, small, torch.bfloat16, device="cuda", zero=True)
                ops.k2s2_gather(x, packed, None, a, out, 1, None)
                return out.buf.clone(), None
        return run

    cases = [("fwd32 multi-item", fwd_case(32, (24, 64, 64)), True), ("fwd64", fwd_case(64, (8, 32, 40)), True),
             ("fwd128 small", fwd_case(128, (4, 16, 16)), True),
             ("wgrad2 64", wgrad_case(64, (8, 32, 32)), False), ("wgrad2 32 clustered", wgrad_case(32, (16, 64, 64)), False),
             ("wgrad per-tap 128", wgrad_case(128, (4, 16, 32)), False),
             ("k2 gather", k2_case(False), True), ("k2 scatter", k2_case(True), True),
             ("k2 scatter-accumulate", k2_case(True, True), True)]
    try:
        for name, run, exact in cases:
            _lib.call("msb_set_tile_scheduler", 0)
            ref, ref_s = run()
            _lib.call("msb_set_tile_scheduler", 1)
            for rep in range(3):
                got, got_s = run()
                if exact:
                    assert torch.equal(got, ref), (name, rep)
                else:
                    assert rel(got, ref) <= 1e-5, (name, rep)
                if ref_s is not None:  # per-CTA f32 partial sums group different tiles: last bits of the f64 totals move
                    assert float((got_s - ref_s).abs().max()) <= 1e-6 * float(ref_s.abs().max() + 1), (name, rep)
    finally:
        _lib.call("msb_set_tile_scheduler", 0)


@pytest.mark.parametrize("c,sigmoid_norm,weighted", [(3, False, False), (3, True, True), (20, False, True), (2, False, True)])
def test_dice_loss_constructor_options_match_oracle(c, sigmoid_norm, weighted):
    """DiceLoss(sigmoid_norm=False) = softmax-normalised Dice and DiceLoss(weight=...) = per-class intersection weights
    (dice_loss.py:36-43,64-65), alone and inside MixedLoss (one fused pass with the cross-entropy), values and gradients."""
    from oracle import vnet_oracle as vo
    from medicalseg_b200.models import losses as L
    torch.manual_seed(8)
    logits = torch.randn(2, c, 8, 9, 10) * 2
    labels = torch.randint(0, c, (2, 8, 9, 10), dtype=torch.int32)
    w = (torch.rand(c) + 0.5) if weighted else None
    for mixed in (False, True):
        lo = logits.clone().requires_grad_(True)
        lg = logits.clone().cuda().requires_grad_(True)
        od = vo.DiceLoss(sigmoid_norm=sigmoid_norm, weight=w)
        gd = L.DiceLoss(sigmoid_norm=sigmoid_norm, weight=None if w is None else w.tolist())
        if mixed:
            oll, odice = vo.MixedLoss([vo.CrossEntropyLoss(), od], [1, 1])(lo, labels)
            gll, gdice = L.MixedLoss([L.CrossEntropyLoss(), gd], [1, 1])(lg, labels.cuda())
            ol, gl = sum(oll), sum(gll)
        else:
            ol, odice = od(lo, labels)
            gl, gdice = gd(lg, labels.cuda())
        ol.backward()
        gl.backward()
        assert abs(float(ol) - float(gl)) <= 1e-5, (mixed, float(ol), float(gl))
        assert float(np.abs(np.asarray(gdice) - np.asarray(odice)).max()) <= 1e-5
        assert float((lg.grad.cpu() - lo.grad).abs().max()) <= 1e-5 * float(lo.grad.abs().max()) + 1e-9
