"""Bring-up probe: runs each kernel family against a torch reference in its OWN subprocess (a faulting kernel
cannot poison the rest) and prints one line per check.  Usage on the GPU box:

    python tests/probe_kernels.py            # all checks
    python tests/probe_kernels.py k5_fwd     # names containing the substring
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _imports():
    import torch
    import torch.nn.functional as F
    from medicalseg_b200 import ops, _lib
    from medicalseg_b200.ops import B8
    return torch, F, ops, B8, _lib


def _rel(a, b):
    import torch
    return float((a - b).abs().max() / (b.abs().max() + 1e-12)), float(
        torch.sqrt(((a - b) ** 2).mean()) / (torch.sqrt((b ** 2).mean()) + 1e-12))


@check
def layout_roundtrip():
    torch, F, ops, B8, _ = _imports()
    x = torch.randn(2, 20, 5, 6, 7, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        b = B8.from_ncdhw(x, dt, 24)
        y = b.to_ncdhw(20)
        ref = x if dt == torch.float32 else x.bfloat16().float()
        print("roundtrip", dt, float((y - ref).abs().max()))


@check
def bn_act():
    torch, F, ops, B8, _ = _imports()
    torch.manual_seed(0)
    n, c, dims = 2, 16, (6, 7, 9)
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 3e-2)):
        y = torch.randn(n, c, *dims, device="cuda") * 2 + 0.5
        r = torch.randn(n, c, *dims, device="cuda")
        gamma = torch.rand(c, device="cuda") + 0.5
        beta = torch.randn(c, device="cuda")
        a1 = torch.rand(c, device="cuda") * 0.5
        a2 = torch.rand(c, device="cuda") * 0.5
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        yb, rb = B8.from_ncdhw(y, dt), B8.from_ncdhw(r, dt)
        yq, rq = yb.to_ncdhw(), rb.to_ncdhw()
        sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
        ops.bn_stats(yb, 1, sums)
        bnbuf = torch.empty(4 * c, device="cuda")
        ops.bn_finalize(sums, n * yb.s, gamma, beta, rm, rv, 0.9, 1e-5, True, c, 1, bnbuf)
        out = B8(n, c, dims, dt, device="cuda")
        ops.bn_act_fwd(yb, out, rb, None, 0, bnbuf, a1, a2, 1)
        # torch reference
        yq_ = yq.clone().requires_grad_(True)
        rq_ = rq.clone().requires_grad_(True)
        g_, b_, a1_, a2_ = [t.clone().requires_grad_(True) for t in (gamma, beta, a1, a2)]
        mean = yq_.mean((0, 2, 3, 4), keepdim=True)
        var = yq_.var((0, 2, 3, 4), unbiased=False, keepdim=True)
        t = (yq_ - mean) * torch.rsqrt(var + 1e-5) * g_.view(1, -1, 1, 1, 1) + b_.view(1, -1, 1, 1, 1)
        act1 = torch.where(t > 0, t, a1_.view(1, -1, 1, 1, 1) * t)
        t2 = act1 + rq_
        ref = torch.where(t2 > 0, t2, a2_.view(1, -1, 1, 1, 1) * t2)
        print("bn_act fwd", dt, _rel(out.to_ncdhw(), ref.detach()), "running", float((rm - 0.1 * mean.flatten()).abs().max()),
              float((rv - (0.9 + 0.1 * var.flatten())).abs().max()))
        go = torch.randn_like(ref)
        ref.backward(go)
        gob = B8.from_ncdhw(go, dt)
        goq = gob.to_ncdhw()
        red = torch.zeros(4 * c, dtype=torch.float64, device="cuda")
        ops.bn_act_bwd_reduce(yb, rb, None, 0, gob, bnbuf, a1, a2, 1, red)
        dy, dres = B8(n, c, dims, dt, device="cuda"), B8(n, c, dims, dt, device="cuda")
        dg, db, da1, da2 = [torch.zeros(c, device="cuda") for _ in range(4)]
        ops.bn_act_bwd_apply(yb, rb, None, 0, gob, bnbuf, a1, a2, red, n * yb.s, True, dy, dres, False, dg, db, da1,
                             da2, 1)
        print("bn_act bwd", dt, "dy", _rel(dy.to_ncdhw(), yq_.grad), "dres", _rel(dres.to_ncdhw(), rq_.grad), "dgamma",
              _rel(dg, g_.grad), "dbeta", _rel(db, b_.grad), "da1", _rel(da1, a1_.grad), "da2", _rel(da2, a2_.grad))


@check
def conv_in():
    torch, F, ops, B8, _ = _imports()
    torch.manual_seed(0)
    n, dims = 2, (9, 11, 37)
    x = torch.rand(n, 1, *dims, device="cuda")
    w = torch.randn(16, 1, 5, 5, 5, device="cuda") * 0.1
    b = torch.randn(16, device="cuda")
    ref = F.conv3d(x, w, b, padding=2)
    for dt in (torch.float32, torch.bfloat16):
        out = B8(n, 16, dims, dt, device="cuda")
        sums = torch.zeros(32, dtype=torch.float64, device="cuda")
        ops.conv_in_fwd(x, w, b, out, 1, sums)
        o = out.to_ncdhw()
        print("conv_in fwd", dt, _rel(o, ref), "sums", float((sums[:16] - o.double().sum((0, 2, 3, 4))).abs().max()),
              float((sums[16:] - (o.double() ** 2).sum((0, 2, 3, 4))).abs().max()))
        dy = torch.randn(n, 16, *dims, device="cuda")
        dyb = B8.from_ncdhw(dy, dt)
        dyq = dyb.to_ncdhw()
        dw, dbias = torch.zeros_like(w), torch.zeros_like(b)
        ops.conv_in_wgrad(x, dyb, dw, dbias)
        dw_ref = torch.nn.grad.conv3d_weight(x, w.shape, dyq, padding=2)
        print("conv_in wgrad", dt, _rel(dw, dw_ref), _rel(dbias, dyq.sum((0, 2, 3, 4))))


@check
def conv_strided():
    torch, F, ops, B8, _ = _imports()
    torch.manual_seed(0)
    cases = [((2, 2, 2), (2, 2, 2), 16, 32, (8, 10, 12)), ((2, 2, 4), (2, 2, 1), 16, 32, (8, 8, 12)),
             ((2, 2, 2), (2, 2, 1), 32, 64, (6, 8, 9))]
    for k, s, ci, co, dims in cases:
        for dt in (torch.float32, torch.bfloat16):
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
            e_f = _rel(out.to_ncdhw(), ref.detach())
            dy = torch.randn_like(ref)
            dyb = B8.from_ncdhw(dy, dt)
            dyq = dyb.to_ncdhw()
            ref.backward(dyq)
            dx = B8(n, ci, dims, dt, device="cuda")
            ops.conv_strided_bwd_data(dyb, w, None, dx, k, s, (0, 0, 0), False, 1, None)
            dw, dbias = torch.zeros_like(w), torch.zeros_like(b)
            ops.conv_strided_wgrad(xb, dyb, dw, dbias, k, s, (0, 0, 0), False)
            print("strided conv", k, s, dt, "fwd", e_f, "dgrad", _rel(dx.to_ncdhw(), xq.grad), "wgrad",
                  _rel(dw, w_.grad), "dbias", _rel(dbias, dyq.sum((0, 2, 3, 4))))
            # transposed conv with weight [ci_T = co, co_T = ci, k]
            wt = torch.randn(co, ci, *k, device="cuda") * 0.1
            bt = torch.randn(ci, device="cuda")
            xs = torch.randn(n, co, *od, device="cuda")
            xsb = B8.from_ncdhw(xs, dt)
            xsq = xsb.to_ncdhw().requires_grad_(True)
            wt_ = wt.clone().requires_grad_(True)
            reft = F.conv_transpose3d(xsq, wt_, bt, stride=s)
            outt = B8(n, ci, reft.shape[2:], dt, device="cuda")
            sums = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
            ops.conv_strided_bwd_data(xsb, wt, bt, outt, k, s, (0, 0, 0), False, 1, sums)
            e_tf = _rel(outt.to_ncdhw(), reft.detach())
            dyt = torch.randn_like(reft)
            dytb = B8.from_ncdhw(dyt, dt)
            dytq = dytb.to_ncdhw()
            reft.backward(dytq)
            dxs = B8(n, co, od, dt, device="cuda")
            ops.conv_strided_fwd(dytb, wt, None, dxs, k, s, (0, 0, 0), 1, None)
            dwt, dbt = torch.zeros_like(wt), torch.zeros_like(bt)
            ops.conv_strided_wgrad(dytb, xsb, dwt, dbt, k, s, (0, 0, 0), True)
            print("   transposed", "fwd", e_tf, "dgrad", _rel(dxs.to_ncdhw(), xsq.grad), "wgrad", _rel(dwt, wt_.grad),
                  "dbias", _rel(dbt, dytq.sum((0, 2, 3, 4))), "sum0",
                  float((sums[:ci] - outt.to_ncdhw().double().sum((0, 2, 3, 4))).abs().max()))


def _k5_case(cin, cout, dims, n=2, swap=0, out_c=None, accumulate=False, ts=0):
    torch, F, ops, B8, _lib = _imports()
    torch.manual_seed(0)
    _lib.call("msb_debug_set", 0, swap)
    _lib.call("msb_debug_set", 3, ts)
    x = torch.randn(n, cin, *dims, device="cuda")
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * (2.0 / (cin * 125)) ** 0.5
    b = torch.randn(cout, device="cuda")
    xb = B8.from_ncdhw(x, torch.bfloat16)
    xq = xb.to_ncdhw()
    wq = w.bfloat16().float()
    ref = F.conv3d(xq, wq, b, padding=2)
    oc = out_c or ((cout + 7) // 8 * 8)
    cout_pad = ops.k5_out_pad(oc)
    packed = torch.empty(ops.k5_packed_bytes(cin, cout_pad), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 0, cin, cout_pad)
    out = B8(n, oc, dims, torch.bfloat16, device="cuda", zero=True)
    sums = torch.zeros(2 * oc, dtype=torch.float64, device="cuda")
    ops.k5_fwd(xb, packed, b, cout, out, False, None, 1, sums)
    torch.cuda.synchronize()
    o = out.to_ncdhw(cout)
    e = _rel(o, ref)
    se = float((sums[:cout] - o.double().sum((0, 2, 3, 4))).abs().max() / (o.double().sum((0, 2, 3, 4)).abs().max() + 1e-9))
    return e, se


@check
def k5_fwd_swap0():
    for cin, cout, dims in ((32, 32, (6, 16, 8)), (16, 16, (5, 20, 11)), (64, 64, (4, 16, 16)), (128, 128, (3, 16, 8)),
                            (256, 256, (2, 8, 8)), (32, 2, (6, 18, 10))):
        oc = 16 if cout == 2 else None
        print("k5 fwd", cin, cout, dims, _k5_case(cin, cout, dims, swap=0, out_c=oc))


@check
def k5_fwd_ts():
    for cin, cout, dims in ((32, 32, (6, 16, 8)), (16, 16, (5, 20, 11)), (64, 64, (7, 16, 16)), (128, 128, (5, 16, 8)),
                            (32, 2, (6, 18, 10)), (32, 32, (9, 7, 13))):
        oc = 16 if cout == 2 else None
        print("k5 fwd unstacked", cin, cout, dims, _k5_case(cin, cout, dims, swap=0, out_c=oc, ts=1))


def _k5_wgrad_case(cin, cout, dims, n=2, swap=0, dy_c=None, v1=0):
    torch, F, ops, B8, _lib = _imports()
    torch.manual_seed(0)
    _lib.call("msb_debug_set", 1, swap)
    _lib.call("msb_debug_set", 2, v1)
    x = torch.randn(n, cin, *dims, device="cuda")
    dyc = dy_c or ((cout + 7) // 8 * 8)
    dy = torch.zeros(n, dyc, *dims, device="cuda")
    dy[:, :cout] = torch.randn(n, cout, *dims, device="cuda")
    xb, dyb = B8.from_ncdhw(x, torch.bfloat16), B8.from_ncdhw(dy, torch.bfloat16)
    xq, dyq = xb.to_ncdhw(), dyb.to_ncdhw(cout)
    ref = torch.nn.grad.conv3d_weight(xq, (cout, cin, 5, 5, 5), dyq, padding=2)
    dw = torch.zeros(cout, cin, 5, 5, 5, device="cuda")
    db = torch.zeros(cout, device="cuda")
    ws = torch.empty(ops.k5_wgrad_workspace_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    ops.k5_wgrad(xb, dyb, dw, db, cout, cin, ws)
    torch.cuda.synchronize()
    return _rel(dw, ref), _rel(db, dyq.sum((0, 2, 3, 4)))


@check
def k5_wgrad_swap0():
    for cin, cout, dims in ((32, 32, (6, 16, 16)), (64, 64, (4, 9, 20)), (128, 128, (3, 8, 16)), (256, 256, (2, 8, 8)),
                            (32, 2, (5, 10, 18)), (16, 16, (5, 8, 16))):
        dyc = 16 if cout == 2 else None
        print("k5 wgrad", cin, cout, dims, _k5_wgrad_case(cin, cout, dims, swap=0, dy_c=dyc))


@check
def k5_wgrad_v1():
    for cin, cout, dims in ((32, 32, (6, 16, 16)), (32, 2, (5, 10, 18)), (16, 16, (5, 8, 16))):
        dyc = 16 if cout == 2 else None
        print("k5 wgrad v1", cin, cout, dims, _k5_wgrad_case(cin, cout, dims, swap=0, dy_c=dyc, v1=1))


@check
def k551_folded():
    """w-folded 5x5x1 path: in_tr (input folded) and out_tr (output folded) vs torch conv3d on the same rounded data"""
    torch, F, ops, B8, _lib = _imports()
    torch.manual_seed(0)
    n = 2
    for dims in ((6, 18, 10), (9, 16, 24), (4, 7, 5)):
        # ---- in_tr: 1 -> 16
        x = torch.rand(n, 1, *dims, device="cuda")
        w = torch.randn(16, 1, 5, 5, 5, device="cuda") * (2.0 / 125) ** 0.5
        b = torch.randn(16, device="cuda")
        xq, wq = x.bfloat16().float(), w.bfloat16().float()
        ref = F.conv3d(xq, wq, b, padding=2)
        xf = B8(n, 16, dims, torch.bfloat16, device="cuda")
        ops.fold_w_f32(x, 1, xf, 1)
        packed = torch.empty(ops.k551_packed_bytes(16, 16), dtype=torch.uint8, device="cuda")
        ops.k551_pack(w, packed, 16, 1, 0, 0, 16, 16)
        y = B8(n, 16, dims, torch.bfloat16, device="cuda", zero=True)
        sums = torch.zeros(32, dtype=torch.float64, device="cuda")
        ops.k551_fwd(xf, packed, b, 16, y, False, None, 1, sums)
        yo = y.to_ncdhw()
        dy = torch.randn(n, 16, *dims, device="cuda")
        dyb = B8.from_ncdhw(dy, torch.bfloat16)
        refw = torch.nn.grad.conv3d_weight(xq, (16, 1, 5, 5, 5), dyb.to_ncdhw(), padding=2)
        dw = torch.zeros_like(w)
        ws = torch.empty(ops.k551_wgrad_workspace_bytes(1, 16, 0), dtype=torch.uint8, device="cuda")
        ops.k551_wgrad(xf, dyb, dw, 16, 1, 0, ws)
        print("k551 in_tr", dims, "fwd", _rel(yo, ref), "sums",
              float((sums[:16] - yo.double().sum((0, 2, 3, 4))).abs().max()), "wgrad", _rel(dw, refw))
        # ---- out_tr: 32 -> C
        for c in (2, 3):
            x = torch.randn(n, 32, *dims, device="cuda")
            w = torch.randn(c, 32, 5, 5, 5, device="cuda") * (2.0 / 4000) ** 0.5
            b = torch.randn(c, device="cuda")
            xb = B8.from_ncdhw(x, torch.bfloat16)
            xq, wq = xb.to_ncdhw(), w.bfloat16().float()
            ref = F.conv3d(xq, wq, b, padding=2)
            packed = torch.empty(ops.k551_packed_bytes(32, 16), dtype=torch.uint8, device="cuda")
            ops.k551_pack(w, packed, c, 32, 0, 1, 32, 16)
            pf = B8(n, 16, dims, torch.float32, device="cuda")
            ops.k551_fwd(xb, packed, None, 5 * c, pf, False, None, 1, None)
            y = B8(n, 8, dims, torch.bfloat16, device="cuda")
            sums = torch.zeros(16, dtype=torch.float64, device="cuda")
            ops.unfold_w(pf, b, c, y, 1, sums)
            yo = y.to_ncdhw(c)
            pad_zero = float(y.to_ncdhw(8)[:, c:].abs().max())
            dy = torch.zeros(n, 8, *dims, device="cuda")
            dy[:, :c] = torch.randn(n, c, *dims, device="cuda")
            dyb = B8.from_ncdhw(dy, torch.bfloat16)
            dyq = dyb.to_ncdhw(c)
            refx = torch.nn.grad.conv3d_input(xq.shape, wq, dyq, padding=2)
            refw = torch.nn.grad.conv3d_weight(xq, (c, 32, 5, 5, 5), dyq, padding=2)
            dpf = B8(n, 16, dims, torch.bfloat16, device="cuda")
            ops.fold_w(dyb, c, dpf, -1)
            packed_b = torch.empty(ops.k551_packed_bytes(16, 32), dtype=torch.uint8, device="cuda")
            ops.k551_pack(w, packed_b, c, 32, 1, 1, 16, 32)
            dx = B8(n, 32, dims, torch.bfloat16, device="cuda")
            ops.k551_fwd(dpf, packed_b, None, 32, dx, False, None, 1, None)
            dw = torch.zeros_like(w)
            ws = torch.empty(ops.k551_wgrad_workspace_bytes(32, c, 1), dtype=torch.uint8, device="cuda")
            ops.k551_wgrad(xb, dpf, dw, c, 32, 1, ws)
            db = torch.zeros(c, device="cuda")
            ops.channel_sum(dyb, c, db)
            print("k551 out_tr C=%d" % c, dims, "fwd", _rel(yo, ref), "pad", pad_zero, "sums",
                  float((sums[:c] - yo.double().sum((0, 2, 3, 4))).abs().max()), "dgrad", _rel(dx.to_ncdhw(), refx),
                  "wgrad", _rel(dw, refw), "dbias", _rel(db, dyq.sum((0, 2, 3, 4))))


@check
def timing_folded():
    torch, F, ops, B8, _lib = _imports()

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n, dims = 2, (128, 128, 128)
    x1 = torch.rand(n, 1, *dims, device="cuda")
    xf = B8(n, 16, dims, torch.bfloat16, device="cuda")
    w = torch.randn(16, 1, 5, 5, 5, device="cuda")
    packed = torch.empty(ops.k551_packed_bytes(16, 16), dtype=torch.uint8, device="cuda")
    ops.k551_pack(w, packed, 16, 1, 0, 0, 16, 16)
    y = B8(n, 16, dims, torch.bfloat16, device="cuda"); y.buf.normal_()
    sums = torch.zeros(32, dtype=torch.float64, device="cuda")
    dw = torch.zeros_like(w)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    print("in_tr folded: fold %.3f ms  conv551 %.3f ms  wgrad551 %.3f ms" % (
        timeit(lambda: ops.fold_w_f32(x1, 1, xf, 1)),
        timeit(lambda: ops.k551_fwd(xf, packed, None, 16, y, False, None, 1, sums)),
        timeit(lambda: ops.k551_wgrad(xf, y, dw, 16, 1, 0, ws))))
    c = 2
    x = B8(n, 32, dims, torch.bfloat16, device="cuda"); x.buf.normal_()
    w = torch.randn(c, 32, 5, 5, 5, device="cuda")
    pk = torch.empty(ops.k551_packed_bytes(32, 16), dtype=torch.uint8, device="cuda")
    ops.k551_pack(w, pk, c, 32, 0, 1, 32, 16)
    pkb = torch.empty(ops.k551_packed_bytes(16, 32), dtype=torch.uint8, device="cuda")
    ops.k551_pack(w, pkb, c, 32, 1, 1, 16, 32)
    pf = B8(n, 16, dims, torch.float32, device="cuda")
    y8 = B8(n, 8, dims, torch.bfloat16, device="cuda"); y8.buf.normal_()
    dpf = B8(n, 16, dims, torch.bfloat16, device="cuda")
    dx = B8(n, 32, dims, torch.bfloat16, device="cuda")
    dw = torch.zeros_like(w)
    s8 = torch.zeros(16, dtype=torch.float64, device="cuda")
    print("out_tr folded: conv551 %.3f ms  unfold %.3f ms  fold(dy) %.3f ms  dgrad551 %.3f ms  wgrad551 %.3f ms" % (
        timeit(lambda: ops.k551_fwd(x, pk, None, 5 * c, pf, False, None, 1, None)),
        timeit(lambda: ops.unfold_w(pf, None, c, y8, 1, s8)),
        timeit(lambda: ops.fold_w(y8, c, dpf, -1)),
        timeit(lambda: ops.k551_fwd(dpf, pkb, None, 32, dx, False, None, 1, None)),
        timeit(lambda: ops.k551_wgrad(x, dpf, dw, c, 32, 1, ws))))


@check
def k2s2_umma():
    """tensor-core 2x2x2/stride-2 gather (Conv3D fwd, ConvT dgrad) and scatter (ConvT fwd, Conv3D dgrad) vs torch"""
    torch, F, ops, B8, _lib = _imports()
    torch.manual_seed(0)
    n = 2
    for ci, co, bd in ((16, 32, (8, 32, 16)), (32, 64, (4, 36, 20)), (64, 128, (6, 8, 8)), (128, 256, (2, 4, 16)),
                       (256, 128, (4, 8, 8)), (64, 16, (8, 32, 16))):
        sd = tuple(d // 2 for d in bd)
        # ---- gather: Conv3D(ci -> co) forward on the big grid
        x = torch.randn(n, ci, *bd, device="cuda")
        w = torch.randn(co, ci, 2, 2, 2, device="cuda") * (2.0 / (ci * 8)) ** 0.5
        b = torch.randn(co, device="cuda")
        xb = B8.from_ncdhw(x, torch.bfloat16)
        xq, wq = xb.to_ncdhw(), w.bfloat16().float()
        ref = F.conv3d(xq, wq, b, stride=2)
        pk = torch.empty(ops.k2s2_packed_bytes(ci, (co + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
        ops.k2s2_pack(w, pk, ci, co, 0, ci, (co + 15) // 16 * 16)
        out = B8(n, co, sd, torch.bfloat16, device="cuda", zero=True)
        sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        ops.k2s2_gather(xb, pk, b, co, out, 1, sums)
        o = out.to_ncdhw()
        e_g = _rel(o, ref)
        e_s = float((sums[:co] - o.double().sum((0, 2, 3, 4))).abs().max() / (o.double().sum((0, 2, 3, 4)).abs().max() + 1e-9))
        # ---- scatter: input gradient of the same conv (dy on the small grid -> big grid), accumulating
        dy = torch.randn(n, co, *sd, device="cuda")
        dyb = B8.from_ncdhw(dy, torch.bfloat16)
        refx = torch.nn.grad.conv3d_input(xq.shape, wq, dyb.to_ncdhw(), stride=2)
        pk1 = torch.empty(ops.k2s2_packed_bytes(co, (ci + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
        ops.k2s2_pack(w, pk1, co, ci, 1, co, (ci + 15) // 16 * 16)
        base = torch.randn(n, ci, *bd, device="cuda")
        dx = B8.from_ncdhw(base, torch.bfloat16)
        baseq = dx.to_ncdhw()
        ops.k2s2_scatter(dyb, pk1, None, ci, dx, True, 1, None)
        e_d = _rel(dx.to_ncdhw(), baseq + refx)
        # ---- scatter as ConvTranspose3d forward (weight [co][ci]: co -> ci channels) with bias and BN sums
        wt = torch.randn(co, ci, 2, 2, 2, device="cuda") * (2.0 / co) ** 0.5
        bt = torch.randn(ci, device="cuda")
        reft = F.conv_transpose3d(dyb.to_ncdhw(), wt.bfloat16().float(), bt, stride=2)
        pk2 = torch.empty(ops.k2s2_packed_bytes(co, (ci + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
        ops.k2s2_pack(wt, pk2, co, ci, 1, co, (ci + 15) // 16 * 16)
        outt = B8(n, ci, bd, torch.bfloat16, device="cuda", zero=True)
        sums2 = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
        ops.k2s2_scatter(dyb, pk2, bt, ci, outt, False, 1, sums2)
        ot = outt.to_ncdhw()
        e_t = _rel(ot, reft)
        e_s2 = float((sums2[:ci] - ot.double().sum((0, 2, 3, 4))).abs().max() / (ot.double().sum((0, 2, 3, 4)).abs().max() + 1e-9))
        print("k2s2", ci, co, bd, "gather", e_g, "sums", e_s, "scatter+acc", e_d, "convT", e_t, "sums", e_s2)


@check
def timing_k2s2():
    torch, F, ops, B8, _lib = _imports()

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n = 2
    for ci, co, d in ((16, 32, 128), (32, 64, 64), (64, 128, 32), (128, 256, 16)):
        x = B8(n, ci, (d,) * 3, torch.bfloat16, device="cuda"); x.buf.normal_()
        y = B8(n, co, (d // 2,) * 3, torch.bfloat16, device="cuda"); y.buf.normal_()
        w = torch.randn(co, ci, 2, 2, 2, device="cuda") * 0.1
        pk = torch.empty(ops.k2s2_packed_bytes(ci, co), dtype=torch.uint8, device="cuda")
        ops.k2s2_pack(w, pk, ci, co, 0, ci, co)
        pk1 = torch.empty(ops.k2s2_packed_bytes(co, ci), dtype=torch.uint8, device="cuda")
        ops.k2s2_pack(w, pk1, co, ci, 1, co, ci)
        sums = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        mb = (x.buf.numel() + y.buf.numel()) * 2 / 1e6
        t_g = timeit(lambda: ops.k2s2_gather(x, pk, None, co, y, 1, sums))
        t_s = timeit(lambda: ops.k2s2_scatter(y, pk1, None, ci, x, False, 1, None))
        print("k2s2 umma %3d<->%3d @%3d: gather %.3f ms (%.0f GB/s)  scatter %.3f ms (%.0f GB/s)" %
              (ci, co, d, t_g, mb / t_g, t_s, mb / t_s))


@check
def prof_fwd():
    """where the MMA-issuing warp of the 5x5x5 forward kernel waits (debug flag 5), per layer shape"""
    import ctypes as C
    torch, F, ops, B8, _lib = _imports()
    n = 2
    for cin, cout, d in ((32, 32, 128), (64, 64, 64), (128, 128, 32), (256, 256, 16)):
        dims = (d,) * 3
        x = B8(n, cin, dims, torch.bfloat16, device="cuda"); x.buf.normal_()
        y = B8(n, cout, dims, torch.bfloat16, device="cuda")
        w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.02
        cp = ops.k5_out_pad(cout)
        packed = torch.empty(ops.k5_packed_bytes(cin, cp), dtype=torch.uint8, device="cuda")
        ops.k5_pack(w, packed, cout, cin, 0, cin, cp)
        sums = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        ops.k5_fwd(x, packed, None, cout, y, False, None, 1, sums)
        _lib.call("msb_debug_set", 5, 1)
        ops.k5_fwd(x, packed, None, cout, y, False, None, 1, sums)
        torch.cuda.synchronize()
        buf = (C.c_longlong * (148 * 4))()
        _lib.call("msb_debug_read_prof", C.cast(buf, C.c_void_p))
        _lib.call("msb_debug_set", 5, 0)
        t = torch.tensor(list(buf), dtype=torch.float64).view(148, 4)
        t = t[t[:, 0] > 0]
        m = t.mean(0)
        print("fwd %d->%d @%d: CTAs %d  MMA-warp clocks: total %.0f  wait weights %.1f%%  wait halo %.1f%%  wait accumulator %.1f%%"
              % (cin, cout, d, t.shape[0], m[0], 100 * m[1] / m[0], 100 * m[2] / m[0], 100 * m[3] / m[0]))


@check
def timing():
    """device time of the dominant layers at the benchmark shapes (batch 2)"""
    torch, F, ops, B8, _lib = _imports()

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n = 2
    for cin, cout, dims, oc in ((32, 32, (128,) * 3, 32), (32, 2, (128,) * 3, 16), (16, 32, (128,) * 3, 32),
                                (64, 64, (64,) * 3, 64), (128, 128, (32,) * 3, 128), (256, 256, (16,) * 3, 256),
                                (256, 256, (8,) * 3, 256)):
        x = B8(n, cin, dims, torch.bfloat16, device="cuda"); x.buf.normal_()
        y = B8(n, oc, dims, torch.bfloat16, device="cuda"); y.buf.normal_()
        w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.02
        cp = ops.k5_out_pad(oc)
        packed = torch.empty(ops.k5_packed_bytes(cin, cp), dtype=torch.uint8, device="cuda")
        ops.k5_pack(w, packed, cout, cin, 0, cin, cp)
        sums = torch.zeros(2 * oc, dtype=torch.float64, device="cuda")
        gf = 2 * n * 125 * cin * cout * dims[0] * dims[1] * dims[2] / 1e9
        res = {}
        for ts in (0, 1):
            if ts == 1 and cp > 64:
                continue
            _lib.call("msb_debug_set", 3, ts)
            res["fwd_stacked" if ts == 0 else "fwd_plain"] = timeit(lambda: ops.k5_fwd(x, packed, None, cout, y, False, None, 1, sums))
        _lib.call("msb_debug_set", 3, 0)
        dw = torch.zeros(cout, cin, 5, 5, 5, device="cuda")
        ws = torch.empty(ops.k5_wgrad_workspace_bytes(cin, cout), dtype=torch.uint8, device="cuda")
        for v1 in (0, 1):
            _lib.call("msb_debug_set", 2, v1)
            res["wgrad_v%d" % (2 - v1)] = timeit(lambda: ops.k5_wgrad(x, y, dw, None, cout, cin, ws))
        _lib.call("msb_debug_set", 2, 0)
        print("k5 %3d->%3d @%s  %.1f GF/pass: " % (cin, cout, dims[0], gf) +
              "  ".join("%s %.3f ms (%.0f TF/s)" % (k, v, gf / v) for k, v in res.items()))
    # down / up convs of the lung config
    for ci, co, d in ((16, 32, 128), (32, 64, 64), (64, 128, 32), (128, 256, 16)):
        x = B8(n, ci, (d,) * 3, torch.bfloat16, device="cuda"); x.buf.normal_()
        y = B8(n, co, (d // 2,) * 3, torch.bfloat16, device="cuda"); y.buf.normal_()
        w = torch.randn(co, ci, 2, 2, 2, device="cuda") * 0.1
        dw = torch.zeros_like(w)
        k = s = (2, 2, 2)
        t_f = timeit(lambda: ops.conv_strided_fwd(x, w, None, y, k, s, (0, 0, 0), 1, None))
        t_d = timeit(lambda: ops.conv_strided_bwd_data(y, w, None, x, k, s, (0, 0, 0), False, 1, None))
        t_w = timeit(lambda: ops.conv_strided_wgrad(x, y, dw, None, k, s, (0, 0, 0), False))
        wsb = torch.empty(ops.k2s2_wgrad_workspace_bytes(n, ci, co, (d,) * 3), dtype=torch.uint8, device="cuda")
        t_w2 = timeit(lambda: ops.k2s2_wgrad(x, y, dw, None, False, wsb))
        print("k2s2 %3d->%3d @%3d: gather %.3f ms  scatter %.3f ms  wgrad %.3f ms  wgrad(tensor core) %.3f ms" %
              (ci, co, d, t_f, t_d, t_w, t_w2))
    x = torch.rand(n, 1, 128, 128, 128, device="cuda")
    y = B8(n, 16, (128,) * 3, torch.bfloat16, device="cuda"); y.buf.normal_()
    w = torch.randn(16, 1, 5, 5, 5, device="cuda")
    dw = torch.zeros_like(w)
    print("in_tr fwd %.3f ms  wgrad %.3f ms" % (timeit(lambda: ops.conv_in_fwd(x, w, None, y, 1, None)),
                                                timeit(lambda: ops.conv_in_wgrad(x, y, dw, None))))


@check
def k5_dgrad():
    torch, F, ops, B8, _lib = _imports()
    torch.manual_seed(0)
    n, cin, cout, dims = 2, 32, 32, (6, 16, 8)
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.02
    dy = torch.randn(n, cout, *dims, device="cuda")
    dyb = B8.from_ncdhw(dy, torch.bfloat16)
    ref = torch.nn.grad.conv3d_input((n, cin, *dims), w.bfloat16().float(), dyb.to_ncdhw(), padding=2)
    packed = torch.empty(ops.k5_packed_bytes(cout, ops.k5_out_pad(cin)), dtype=torch.uint8, device="cuda")
    ops.k5_pack(w, packed, cout, cin, 1, cout, ops.k5_out_pad(cin))
    base = torch.randn(n, cin, *dims, device="cuda")
    dx = B8.from_ncdhw(base, torch.bfloat16)
    baseq = dx.to_ncdhw()
    scale = (torch.rand(n, cin, device="cuda") > 0.5).float() * 2
    ops.k5_fwd(dyb, packed, None, cin, dx, True, scale, 1, None)
    print("k5 dgrad accumulate+scale", _rel(dx.to_ncdhw(), baseq + ref * scale.view(n, cin, 1, 1, 1)))


@check
def loss():
    torch, F, ops, B8, _ = _imports()
    from oracle import vnet_oracle as vo
    from medicalseg_b200.models import losses as L
    torch.manual_seed(0)
    for c in (2, 3, 20):
        logits = torch.randn(2, c, 8, 9, 10) * 2
        labels = torch.randint(0, c, (2, 8, 9, 10), dtype=torch.int32)
        lo = logits.clone().requires_grad_(True)
        ol = vo.default_losses()
        ll, dice = vo.loss_computation([lo], labels, ol)
        sum(ll).backward()
        lg = logits.cuda().requires_grad_(True)
        ours = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
        l2, d2 = L.loss_computation([lg], labels.cuda(), ours)
        sum(l2).backward()
        print("loss C=%d" % c, [float(a) - float(b) for a, b in zip(ll, l2)], float(abs(dice - d2).max()),
              _rel(lg.grad.cpu(), lo.grad))


@check
def preprocess():
    torch, F, ops, B8, _ = _imports()
    import numpy as np
    from medicalseg_b200 import preprocess as P
    G = np.load(os.path.join(ROOT, "tests/golden/preprocess_ref.npz"))
    for name, shp in (("iso", (16, 16, 16)), ("aniso", (16, 24, 12)), ("up", (20, 17, 23)), ("one", (1, 4, 8))):
        hu = P.HUnorm(G[name + "_in"])
        r1, _ = P.resample(G[name + "_hunorm"], new_shape=list(shp), order=1)
        rf, _ = P.resample(G[name + "_in"], new_shape=list(shp), order=1, pre_op=("hunorm", -1200, 600, -2000))
        r0, _ = P.resample(G[name + "_label"], new_shape=list(shp), order=0)
        nm = P.normalize(np.nan_to_num(G[name + "_in"], nan=0.0))
        nf = P.normalize(np.nan_to_num(G[name + "_in"], nan=0.0), 0, 2650)
        print("preprocess", name, "hunorm", float(np.abs(hu - G[name + "_hunorm"]).max()), "res1",
              float(np.abs(r1 - G[name + "_resample1"]).max()), "fused", float(np.abs(rf - G[name + "_resample1"]).max()),
              "res0", int(np.abs(r0 - G[name + "_resample0"]).max()), "norm",
              float(np.abs(nm - G[name + "_norm_minmax"]).max()), float(np.abs(nf - G[name + "_norm_fixed"]).max()))
    print("remap", int(np.abs(P.label_remap(G["remap_in"], {1: 2, 2: 3, 5: 0}) - G["remap_out"]).max()))


def _vnet_case(dtype, num_classes, shape, ks=None, train=True, steps=1):
    torch, F, ops, B8, _ = _imports()
    from oracle import vnet_oracle as vo
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    torch.manual_seed(0)
    kw = {} if ks is None else dict(kernel_size=ks[0], stride_size=ks[1])
    om = vo.VNetOracle(num_classes=num_classes, **kw)
    img, lab = vo.synthetic_batch(2, shape, num_classes, seed=0)
    m = VNet(num_classes=num_classes, compute_dtype=dtype, **kw)
    m.set_state_dict(om.state_dict())
    ol = vo.default_losses()
    ours = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    oopt = vo.Momentum(vo.PolynomialDecay(0.001, 15000), list(om.parameters()), 0.9, 1e-4)
    opt = Momentum(PolynomialDecay(0.001, 15000), m.parameters(), 0.9, 1e-4)
    if train:
        om.train(); m.train()
    else:
        om.eval(); m.eval()
    for step in range(steps):
        masks = vo.make_dropout_masks(2, seed=0, step=step) if train else None
        ologits = om(img, masks)[0]
        ll, dice = vo.loss_computation([ologits], lab, ol)
        sum(ll).backward()
        m.set_dropout_masks(masks)
        logits = m(img.cuda())[0]
        l2, d2 = L.loss_computation([logits], lab.cuda(), ours)
        sum(l2).backward()
        print("vnet", dtype, num_classes, shape, "step", step, "logits", _rel(logits.detach().cpu(), ologits.detach()),
              "loss", [float(a) for a in ll], [float(b) for b in l2], "dice diff", float(abs(dice - d2).max()))
        osd = dict(om.named_parameters())
        worst = (0, "")
        for name, p in m.named_parameters():
            g = m.store.grad_view(name).cpu()
            og = osd[name].grad
            cos = float((g * og).sum() / (g.norm() * og.norm() + 1e-30))
            rel = float((g - og).norm() / (og.norm() + 1e-30))
            if og.norm() > 1e-7 and rel > worst[0]:
                worst = (rel, name, cos)
        print("   worst grad rel err", worst)
        oopt.step(); oopt.lr.step(); oopt.clear_grad()
        opt.step(); opt._learning_rate.step(); m.clear_gradients()
        pw = max(float((m.store.view(n).cpu() - p.detach()).abs().max()) for n, p in om.named_parameters())
        rm = max(float((m.store.view(n).cpu() - b).abs().max()) for n, b in om.named_buffers())
        print("   max param diff after step", pw, "max running-stat diff", rm)


@check
def vnet_f32_train():
    _vnet_case("f32", 2, (16, 16, 16), steps=2)


@check
def vnet_f32_eval():
    _vnet_case("f32", 3, (16, 16, 16), train=False)


@check
def vnet_f32_mri():
    _vnet_case("f32", 20, (32, 32, 12), ks=([[2, 2, 4], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
                                             [[2, 2, 1], [2, 2, 1], [2, 2, 2], [2, 2, 2]]))


@check
def vnet_bf16_train():
    _vnet_case("bf16", 2, (16, 16, 16), steps=2)


@check
def vnet_bf16_32():
    _vnet_case("bf16", 2, (32, 32, 32), steps=1)


@check
def vnet_bf16_cosines():
    """gradient cosine per conv weight, bf16 path vs f32 oracle, at 32^3 and 64^3"""
    torch, F, ops, B8, _ = _imports()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_vnet as tv
    for shape in ((32, 32, 32), (64, 64, 64)):
        vo, L, om, m, img, lab, ol, ours = tv._setup("bf16", 2, shape, True)
        ologits, logits, ll, l2, dice, d2 = tv._step(vo, L, om, m, img, lab, ol, ours, True)
        worst, cos = tv._grad_errors(om, m)
        rms = float(torch.sqrt(((logits - ologits) ** 2).mean()) / torch.sqrt((ologits ** 2).mean()))
        print(shape, "rms", rms, "dice diff", float(abs(dice - d2).max()), "worst rel", worst)
        print("   ", {k: round(v, 4) for k, v in cos.items()})


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        CHECKS[sys.argv[2]]()
        return
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    for name in CHECKS:
        if pat not in name:
            continue
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], capture_output=True,
                               text=True, timeout=300)
            status = "OK " if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
            print("=== %s %s" % (name, status))
            print(r.stdout.strip())
            if r.returncode != 0:
                print(r.stderr.strip()[-1500:])
        except subprocess.TimeoutExpired:
            print("=== %s TIMEOUT" % name)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
