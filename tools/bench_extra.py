"""Secondary measurements (not the driver's bench line): BASELINE.json configs[3] (MRI 512x512x12, 20 classes,
anisotropic kernels) and configs[4] (preprocess 512^3 -> 128^3), each with its CPU baseline on the box's host cores.

    python tools/bench_extra.py mri        # VNet MRISpineSeg train step, batch 2, bf16
    python tools/bench_extra.py preprocess # HUnorm + resample (order 1) and label resample (order 0)
    python tools/bench_extra.py infer      # eval-mode forward, both configs
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def mri(deepsup=False):
    from medicalseg_b200.models import VNet, VNetDeepSup, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    kw = dict(kernel_size=[[2, 2, 4], [2, 2, 2], [2, 2, 2], [2, 2, 2]], stride_size=[[2, 2, 1], [2, 2, 1], [2, 2, 2], [2, 2, 2]])
    m = (VNetDeepSup if deepsup else VNet)(num_classes=20, compute_dtype="bf16", **kw)
    m.train()
    nl = 4 if deepsup else 1  # vnetdeepsup_mri_spine_seg_512_512_12_15k.yml: four MixedLoss objects x 0.25
    losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1]) for _ in range(nl)],
              "coef": [1.0 / nl] * nl}
    opt = Momentum(PolynomialDecay(0.1, 15000), m.parameters(), 0.9, 1e-4)
    img = torch.rand(2, 1, 512, 512, 12, device="cuda")
    lab = torch.randint(0, 20, (2, 512, 512, 12), device="cuda", dtype=torch.int32)

    def step():
        ll, dice = L.loss_computation(m(img), lab, losses)
        sum(ll).backward()
        opt.step(); opt._learning_rate.step(); m.clear_gradients()

    for _ in range(3):
        step()
    ms = ev_time(step, 5)
    print(json.dumps({"metric": "%s MRISpineSeg 512x512x12 20-class bf16 train-step volumes/sec" % type(m).__name__,
                      "value": round(2e3 / ms, 3),
                      "unit": "volumes/s", "ms_per_step": round(ms, 2), "n_gpus": 1, "batch": 2,
                      "tflops": round(2 * 12817 / ms, 1)}))


def fp32(dtype="f32"):
    """BASELINE configs[2] on one GPU.  'f32': the PARITY path (CUDA-core direct convolutions, f32 storage);
    'f32x3': f32 storage with the 5x5x5 convolutions as three bf16 tensor-core passes"""
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    m = VNet(num_classes=2, compute_dtype=dtype)
    m.train()
    losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    opt = Momentum(PolynomialDecay(0.001, 15000), m.parameters(), 0.9, 1e-4)
    img = torch.rand(2, 1, 128, 128, 128, device="cuda")
    lab = torch.randint(0, 2, (2, 128, 128, 128), device="cuda", dtype=torch.int32)

    def step():
        ll, dice = L.loss_computation(m(img), lab, losses)
        sum(ll).backward()
        opt.step(); opt._learning_rate.step(); m.clear_gradients()

    step()
    ms = ev_time(step, 2)
    print(json.dumps({"metric": "VNet 128^3 fp32 storage (%s) train-step volumes/sec" % dtype, "value": round(2e3 / ms, 3),
                      "unit": "volumes/s", "ms_per_step": round(ms, 1), "n_gpus": 1, "batch": 2,
                      "tflops": round(2 * 4380.9 / ms, 1)}))


def infer():
    """eval-mode forward (core/val.py:101-118, core/infer.py:79-92): logits for a batch of 2 volumes, bf16 engine"""
    from medicalseg_b200.models import VNet
    for name, kw, shape in (("128^3 2-class", {}, (128, 128, 128)),
                            ("MRISpineSeg 512x512x12 20-class",
                             dict(num_classes=20, kernel_size=[[2, 2, 4], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
                                  stride_size=[[2, 2, 1], [2, 2, 1], [2, 2, 2], [2, 2, 2]]), (512, 512, 12))):
        kw.setdefault("num_classes", 2)
        m = VNet(compute_dtype="bf16", **kw)
        m.eval()
        img = torch.rand(2, 1, *shape, device="cuda")
        with torch.no_grad():
            for _ in range(3):
                m(img)
            ms = ev_time(lambda: m(img), 10)
        print(json.dumps({"metric": "VNet %s eval forward volumes/sec" % name, "value": round(2e3 / ms, 2),
                          "unit": "volumes/s", "ms_per_batch": round(ms, 3), "batch": 2, "n_gpus": 1}))
        # one evaluate() step (core/val.py:101-118): prediction + CE/Dice of the same logits
        from medicalseg_b200.models import losses as L
        lab = torch.randint(0, kw["num_classes"], (2, *shape), device="cuda", dtype=torch.int32)
        cfg = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}

        def unfused():
            logits = m(img)[0]
            pred = torch.argmax(logits, dim=1, keepdim=True).to(torch.int32)
            loss, dice = L.loss_computation([logits], lab, cfg)
            return pred, loss

        with torch.no_grad():
            for _ in range(2):
                unfused(); m.predict_with_losses(img, lab, cfg)
            ms_u = ev_time(unfused, 10)
            ms_f = ev_time(lambda: m.predict_with_losses(img, lab, cfg), 10)
        print(json.dumps({"metric": "VNet %s evaluate step (pred + CE/Dice) volumes/sec" % name,
                          "fused_head": round(2e3 / ms_f, 2), "unfused": round(2e3 / ms_u, 2), "unit": "volumes/s",
                          "ms_fused": round(ms_f, 3), "ms_unfused": round(ms_u, 3), "batch": 2}))


def preprocess():
    from medicalseg_b200 import preprocess as P
    import scipy.ndimage
    rng = np.random.default_rng(0)
    vol = rng.uniform(-2000, 2000, size=(512, 512, 512)).astype(np.float32)
    vol[rng.random(vol.shape) < 0.001] = np.nan
    lab = rng.integers(0, 3, size=(512, 512, 512)).astype(np.int32)
    dvol, dlab = torch.from_numpy(vol).cuda(), torch.from_numpy(lab).cuda()
    # device-resident: fused HUnorm+resample, and the reference's two-pass order
    t_fused = ev_time(lambda: P.resample(dvol, new_shape=[128, 128, 128], order=1, pre_op=("hunorm", -1200, 600, -2000)), 10)
    t_two = ev_time(lambda: P.resample(P.HUnorm(dvol), new_shape=[128, 128, 128], order=1), 10)
    t_lab = ev_time(lambda: P.resample(dlab, new_shape=[128, 128, 128], order=0), 10)
    # host-resident through the public API (numpy in, numpy out; H2D + D2H inside)
    hv = torch.from_numpy(vol).pin_memory()
    t0 = time.perf_counter()
    for _ in range(3):
        out, _ = P.resample(hv.cuda(non_blocking=True), new_shape=[128, 128, 128], order=1,
                            pre_op=("hunorm", -1200, 600, -2000))
        res = out.cpu()
    t_host = (time.perf_counter() - t0) / 3 * 1e3
    # CPU baseline: what the reference executes (numpy HUnorm + scipy.ndimage.zoom, single-threaded C)
    t0 = time.perf_counter()
    # tools/preprocess_utils/values.py:67-87 restated inline: nan_to_num(nan=-2000) -> window [-1200, 600] -> [0, 255]
    hu = np.clip((np.nan_to_num(vol, nan=-2000.0) - (-1200.0)) / ((600.0 - (-1200.0)) / 255.0), 0, 255)
    ref = scipy.ndimage.zoom(hu.astype(np.float32), 0.25, mode="nearest", order=1)
    t_cpu = (time.perf_counter() - t0) * 1e3
    err = float(np.abs(res.numpy() - ref).max())
    full_scan_mb = (512 ** 3 * 4 + 128 ** 3 * 4) / 1e6
    print(json.dumps({"metric": "preprocess HUnorm+resample 512^3->128^3 volumes/sec",
                      "device_fused_ms": round(t_fused, 3), "device_two_pass_ms": round(t_two, 3),
                      "label_order0_ms": round(t_lab, 3), "host_pinned_e2e_ms": round(t_host, 2),
                      "cpu_reference_ms": round(t_cpu, 1), "cpu_cores": 1,
                      "volumes_per_s_device": round(1e3 / t_fused, 1), "volumes_per_s_host": round(1e3 / t_host, 2),
                      "volumes_per_s_cpu": round(1e3 / t_cpu, 3),
                      "two_pass_full_scan_GBps": round((full_scan_mb + 2 * 512 ** 3 * 4 / 1e6) / t_two, 1),
                      "fused_compulsory_GBps": round(142.6 / t_fused, 1), "max_abs_err_vs_scipy": err}))


def augment():
    """the lung_coronavirus.yml training pipeline (RandomResizedCrop3D 128 -> RandomRotation3D 90 -> RandomFlip3D ->
    Compose / max) on one 128^3 sample: device kernels (host .npy already in pinned memory, H2D inside the timed
    region) vs the oracle restatement of the reference's NumPy/SciPy path on one core (what a DataLoader worker runs)"""
    import random
    from medicalseg_b200 import transforms as T
    from oracle import transforms_oracle as to
    rng = np.random.default_rng(0)
    img = (rng.random((128, 128, 128)) * 255).astype(np.float32)
    lab = rng.integers(0, 3, size=(128, 128, 128)).astype(np.int32)
    himg, hlab = torch.from_numpy(img).pin_memory(), torch.from_numpy(lab).pin_memory()
    pipe = T.Compose([T.RandomResizedCrop3D(size=128, scale=[0.8, 1.2]), T.RandomRotation3D(degrees=90),
                      T.RandomFlip3D()])
    random.seed(0)

    def dev_item():
        return pipe(himg.cuda(non_blocking=True), hlab.cuda(non_blocking=True))

    for _ in range(3):
        dev_item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        out = dev_item()
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / n * 1e3
    random.seed(0)
    dimg, dlab = himg.cuda(), hlab.cuda()
    t_kern = ev_time(lambda: pipe(dimg, dlab), 10)
    opipe = to.Compose([to.RandomResizedCrop3D(size=128, scale=[0.8, 1.2]), to.RandomRotation3D(degrees=90),
                        to.RandomFlip3D()])
    random.seed(0)
    t0 = time.perf_counter()
    for _ in range(2):
        opipe(img, lab)
    t_cpu = (time.perf_counter() - t0) / 2 * 1e3
    import scipy.ndimage
    t0 = time.perf_counter()
    scipy.ndimage.rotate(img, angle=33.3, axes=[0, 1], order=1, cval=0, reshape=False)
    scipy.ndimage.rotate(lab, angle=33.3, axes=[0, 1], order=1, cval=0, reshape=False)
    t_scipy_rot = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"metric": "augmentation pipeline (crop+zoom, rotate, flip, /max) 128^3 samples/sec",
                      "device_from_pinned_host_ms": round(t_dev, 3), "device_resident_ms": round(t_kern, 3),
                      "samples_per_s_device": round(1e3 / t_dev, 1),
                      "cpu_oracle_numpy_ms": round(t_cpu, 1), "scipy_rotate_image_and_label_ms": round(t_scipy_rot, 1),
                      "cpu_cores": 1, "train_step_ms_per_sample": 6.0}))


if __name__ == "__main__":
    {"mri": mri, "preprocess": preprocess, "infer": infer, "fp32": fp32, "augment": augment,
     "deepsup": lambda: mri(True), "fp32x3": lambda: fp32("f32x3")}[sys.argv[1]]()
