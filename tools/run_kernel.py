"""Runs ONE kernel family alone a few times (so that `ncu --set full -k regex:<name>` captures exactly it) and prints
its CUDA-event time.  Shapes are the BASELINE configs[1] ones (batch 2).

    python tools/run_kernel.py fwd32      # conv_k5_fwd   32->32 @128^3  (dominant layer, up_tr32.ops[0].conv1)
    python tools/run_kernel.py wgrad32    # conv_k5_wgrad 32->32 @128^3
    python tools/run_kernel.py fwd64 | wgrad64 | fwd128 | wgrad128 | fwd256 | wgrad256 | fwd256s | wgrad256s
    python tools/run_kernel.py splitk256 | splitk256s | splitk128s   # split-K forward + finalize (8^3 / 16^3 levels)
    python tools/run_kernel.py head20 | loss20 | trilinear | preprocess | mrifwd32 | mriwgrad128
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

CASES = {  # name -> (channels, cube edge)
    "32": (32, 128), "64": (64, 64), "128": (128, 32), "256": (256, 16), "256s": (256, 8), "128s": (128, 16),
    "64s": (64, 32), "32s": (32, 64),
}
GFLOP = lambda c, e: 2 * 125 * c * c * e ** 3 / 1e9  # noqa: E731  (per volume, one pass)


def k2s2(what, reps):
    """k2scatter: up_tr32.up_conv fwd (64 -> 16, 64^3 -> 128^3); k2scatter_acc: down_tr32.down_conv dgrad (32 -> 16,
    accumulate); k2gather: down_tr32.down_conv fwd (16 -> 32, 128^3 -> 64^3)"""
    from medicalseg_b200 import ops
    from medicalseg_b200.ops import B8
    dev = torch.device("cuda", 0)
    n, big, small = 2, (128, 128, 128), (64, 64, 64)
    if what == "k2gather":
        a, b = 32, 16            # weight [A][B][2][2][2]: gather produces A from B
        w = torch.randn(a, b, 2, 2, 2, device=dev) * 0.1
        x = B8(n, b, big, torch.bfloat16, device=dev); x.buf.normal_()
        out = B8(n, a, small, torch.bfloat16, device=dev)
        packed = torch.empty(ops.k2s2_packed_bytes(b, 32), dtype=torch.uint8, device=dev)
        ops.k2s2_pack(w, packed, b, a, 0, b, 32)
        fn = lambda: ops.k2s2_gather(x, packed, None, a, out, 1, None)  # noqa: E731
        mb = (x.buf.numel() + out.buf.numel()) * 2 / 1e6
    else:
        a, b = (64, 16) if what == "k2scatter" else (32, 16)
        acc = what == "k2scatter_acc"
        w = torch.randn(a, b, 2, 2, 2, device=dev) * 0.1
        x = B8(n, a, small, torch.bfloat16, device=dev); x.buf.normal_()
        out = B8(n, b, big, torch.bfloat16, device=dev, zero=True)
        packed = torch.empty(ops.k2s2_packed_bytes(a, 16), dtype=torch.uint8, device=dev)
        ops.k2s2_pack(w, packed, a, b, 1, a, 16)
        fn = lambda: ops.k2s2_scatter(x, packed, None, b, out, acc, 1, None)  # noqa: E731
        mb = (x.buf.numel() + out.buf.numel() * (2 if acc else 1)) * 2 / 1e6
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %.4f ms/call, %.1f MB algorithmic -> %.0f GB/s" % (what, ms, mb, mb / ms))


def bn(what, reps):
    """bn32: the BatchNorm + residual + PReLU kernels of up_tr32.ops[0] (32 channels @128^3, batch 2)"""
    from medicalseg_b200 import ops
    from medicalseg_b200.ops import B8
    dev = torch.device("cuda", 0)
    n, c, dims = 2, 32, (128, 128, 128)
    mk = lambda: B8(n, c, dims, torch.bfloat16, device=dev)  # noqa: E731
    y, r, out, go, dy, dres = mk(), mk(), mk(), mk(), mk(), mk()
    for t in (y, r, go):
        t.buf.normal_()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    a1, a2 = torch.full((c,), 0.25, device=dev), torch.full((c,), 0.25, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
    ops.bn_stats(y, 1, sums)
    bnbuf = torch.empty(4 * c, device=dev)
    count = y.s * n
    red = torch.zeros(4 * c, dtype=torch.float64, device=dev)
    grads = [torch.zeros(c, device=dev) for _ in range(4)]
    mb = y.buf.numel() * 2 / 1e6
    cases = {
        "fwd": (lambda: ops.bn_fwd_fused(y, out, r, None, 0, sums, count, gamma, beta, rm, rv, 0.9, 1e-5, True, bnbuf, a1,
                                         a2, 1), 3 * mb),
        "reduce": (lambda: ops.bn_act_bwd_reduce(y, r, None, 0, go, bnbuf, a1, a2, 1, red), 3 * mb),
        "apply": (lambda: ops.bn_act_bwd_apply(y, r, None, 0, go, bnbuf, a1, a2, red, count, True, dy, dres, False,
                                               *grads, 1), 5 * mb),
    }
    cases["fwd"][0]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, (fn, mbytes) in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("bn %s 32ch @128^3 batch 2: %.4f ms, %.0f MB -> %.0f GB/s" % (name, ms, mbytes, mbytes / ms))


def _time(fn, reps, label, unit_work=None, unit="GB/s"):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = "" if unit_work is None else ", %.1f %s" % (unit_work / ms, unit)
    print("%s: %.4f ms/call%s" % (label, ms, extra))
    return ms


def misc(what, reps):
    """kernels DESIGN.md section 4 lists that had CUDA-event timings only: captured here one at a time for ncu"""
    from medicalseg_b200 import ops
    from medicalseg_b200.ops import B8
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if what.startswith("splitk"):
        c, e = {"splitk256": (256, 16), "splitk256s": (256, 8), "splitk128s": (128, 16)}[what]
        n, dims = 2, (e, e, e)
        x = B8(n, c, dims, torch.bfloat16, device=dev); x.buf.normal_()
        y = B8(n, c, dims, torch.bfloat16, device=dev)
        w = torch.randn(c, c, 5, 5, 5, device=dev) * 0.02
        packed = torch.empty(ops.k5_packed_bytes(c, ops.k5_out_pad(c)), dtype=torch.uint8, device=dev)
        ops.k5_pack(w, packed, c, c, 0, c, ops.k5_out_pad(c))
        ws = torch.zeros(max(ops.k5_fwd_workspace_bytes(n, c, dims, c), 16), dtype=torch.uint8, device=dev)
        sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
        bias = torch.zeros(c, device=dev)
        _time(lambda: ops.k5_fwd(x, packed, bias, c, y, False, None, 1, sums, ws), reps,
              "%s: %d->%d 5x5x5 @%d^3 split-K + finalize" % (what, c, c, e), n * GFLOP(c, e), "TFLOP/s")
    elif what in ("mrifwd32", "mriwgrad128", "mriwgrad256"):
        if what == "mrifwd32":
            c, dims = 32, (512, 512, 12)
        elif what == "mriwgrad256":
            c, dims = 256, (64, 64, 4)
        else:
            c, dims = 128, (128, 128, 8)
        n = 2
        x = B8(n, c, dims, torch.bfloat16, device=dev); x.buf.normal_()
        y = B8(n, c, dims, torch.bfloat16, device=dev); y.buf.normal_()
        w = torch.randn(c, c, 5, 5, 5, device=dev) * 0.02
        gf = 2 * 125 * c * c * dims[0] * dims[1] * dims[2] / 1e9
        if what == "mrifwd32":
            packed = torch.empty(ops.k5_packed_bytes(c, ops.k5_out_pad(c)), dtype=torch.uint8, device=dev)
            ops.k5_pack(w, packed, c, c, 0, c, ops.k5_out_pad(c))
            sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
            bias = torch.zeros(c, device=dev)
            _time(lambda: ops.k5_fwd(x, packed, bias, c, y, False, None, 1, sums), reps,
                  "mrifwd32: 32->32 5x5x5 @512x512x12", n * gf, "TFLOP/s")
        else:
            dw = torch.zeros(125 * c * c, device=dev)
            _time(lambda: ops.k5_wgrad_tm(x, y, dw, None, c, c), reps,
                  "%s: %d->%d 5x5x5 wgrad @%s (per-tap kernel)" % (what, c, c, "x".join(map(str, dims))), n * gf, "TFLOP/s")
    elif what in ("head20", "loss20"):
        c, dims, n = 20, (512, 512, 12), 2
        s = dims[0] * dims[1] * dims[2]
        lab = torch.randint(0, c, (n, *dims), device=dev, dtype=torch.int32)
        cw = torch.ones(c, device=dev)
        if what == "head20":
            a = B8(n, 32, dims, torch.bfloat16, device=dev); a.buf.normal_()
            w2, b2 = torch.randn(c, c, device=dev) * 0.3, torch.zeros(c, device=dev)
            pred = torch.empty((n, 1, *dims), dtype=torch.int32, device=dev)
            acc = torch.zeros(3 * c + 2, dtype=torch.float64, device=dev)
            mb = (a.buf.numel() * 2 + 2 * n * s * 4) / 1e6
            _time(lambda: ops.eval_head(a, w2, b2, lab, cw, c, 255, pred=pred, acc=acc), reps,
                  "eval_head C=20 @512x512x12 batch 2 (%.0f MB)" % mb, mb)
        else:
            logits = torch.randn(n, c, *dims, device=dev)
            acc = torch.zeros(3 * c + 2, dtype=torch.float64, device=dev)
            dlog = torch.empty_like(logits)
            mb = (logits.numel() * 4 + n * s * 4) / 1e6
            _time(lambda: ops.dice_ce_fwd(logits, lab, cw, 255, acc), reps, "dice_ce_fwd C=20 (%.0f MB)" % mb, mb)
            _time(lambda: ops.dice_ce_bwd(logits, lab, cw, acc, 255, 1.0, 1.0, None, dlog), reps,
                  "dice_ce_bwd C=20 (%.0f MB)" % (mb + logits.numel() * 4 / 1e6), mb + logits.numel() * 4 / 1e6)
    elif what == "trilinear":
        n, c = 2, 20
        src = torch.randn(n, c, 128, 128, 8, device=dev)
        dst = torch.empty(n, c, 512, 512, 12, device=dev)
        dsrc = torch.empty_like(src)
        mb = (src.numel() + dst.numel()) * 4 / 1e6
        _time(lambda: ops.trilinear_fwd(src, dst), reps, "trilinear_fwd 20ch 128x128x8 -> 512x512x12 (%.0f MB)" % mb, mb)
        _time(lambda: ops.trilinear_bwd(dst, dsrc), reps, "trilinear_bwd (%.0f MB)" % mb, mb)
    elif what == "preprocess":
        from medicalseg_b200 import preprocess as P
        # 4 copies in rotation: the ~140 MB a scan touches must not be served from the 126 MB L2 on the next call
        vols = [torch.empty(512, 512, 512, device=dev).uniform_(-2000, 2000) for _ in range(4)]
        labs = [torch.randint(0, 3, (512, 512, 512), device=dev, dtype=torch.int32) for _ in range(4)]
        k = {"i": 0}

        def nxt(seq):
            k["i"] += 1
            return seq[k["i"] % 4]
        _time(lambda: P.resample(nxt(vols), new_shape=[128, 128, 128], order=1, pre_op=("hunorm", -1200, 600, -2000)),
              max(reps, 8), "resample_f32 + fused HUnorm 512^3 -> 128^3 (142.6 MB compulsory)", 142.6)
        _time(lambda: P.resample(nxt(labs), new_shape=[128, 128, 128], order=0), max(reps, 8), "resample_i32 order 0",
              None)
    else:
        raise SystemExit("unknown case %s" % what)


def main():
    from medicalseg_b200 import ops
    from medicalseg_b200.ops import B8
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd32"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    for key in range(8):  # MSB_DEBUG6=4: clustered kh-stacked wgrad everywhere; MSB_DEBUG0=<permille>: cluster share of
        if os.environ.get("MSB_DEBUG%d" % key):  # the kw-replicated leftover group; MSB_DEBUG6=16: replica form off
            from medicalseg_b200 import _lib
            _lib.call("msb_debug_set", key, int(os.environ["MSB_DEBUG%d" % key]))
    if what.startswith(("splitk", "mri", "head", "loss", "trilinear", "preprocess")):
        return misc(what, reps)
    if what.startswith("k2"):
        return k2s2(what, reps)
    if what.startswith("bn"):
        return bn(what, reps)
    kind = "wgrad" if what.startswith("wgrad") else "fwd"
    c, e = CASES[what[len(kind):]]
    n, dims = 2, (e, e, e)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = B8(n, c, dims, torch.bfloat16, device=dev)
    x.buf.normal_()
    y = B8(n, c, dims, torch.bfloat16, device=dev)
    y.buf.normal_()
    w = torch.randn(c, c, 5, 5, 5, device=dev) * 0.02
    bias = torch.zeros(c, device=dev)
    if kind == "fwd":
        packed = torch.empty(ops.k5_packed_bytes(c, ops.k5_out_pad(c)), dtype=torch.uint8, device=dev)
        ops.k5_pack(w, packed, c, c, 0, c, ops.k5_out_pad(c))
        sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
        fn = lambda: ops.k5_fwd(x, packed, bias, c, y, False, None, 1, sums)  # noqa: E731
    else:
        dw, db = torch.zeros_like(w), torch.zeros(c, device=dev)
        ws = torch.empty(ops.k5_wgrad_workspace_bytes(c, c), dtype=torch.uint8, device=dev)
        fn = lambda: ops.k5_wgrad(x, y, dw, db, c, c, ws)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %d->%d 5x5x5 @%d^3 batch %d: %.4f ms/call (all kernels of the op), %.1f TFLOP/s" %
          (what, c, c, e, n, ms, n * GFLOP(c, e) / ms))


if __name__ == "__main__":
    main()
