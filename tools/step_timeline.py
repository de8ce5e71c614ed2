"""GPU timeline of the bench train step via torch.profiler (CUPTI): per-kernel device time inside real steps
(warm caches, real overlap) plus the idle gaps between kernels - complements the serialized ncu launch list.

    python tools/step_timeline.py [steps] [mode] # prints a per-kernel table for the traced steps
    torchrun --nproc-per-node N tools/step_timeline.py 3   # data-parallel (eager launches): adds the NCCL report - every
                                                           # all-reduce kernel with its start offset / duration and
                                                           # the time the optimizer waited after the last backward kernel
"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    mode = sys.argv[2] if len(sys.argv) > 2 else ""
    mri = mode in ("mri", "evalmri", "deepsup")  # BASELINE configs[3]: 512x512x12, 20 classes, anisotropic
    evaluate = mode in ("eval", "evalmri")  # one evaluate() step: eval-mode forward + fused head (core/val.py:101-118)
    cdt = "f32x3" if len(sys.argv) > 2 and sys.argv[2] == "f32x3" else "bf16"  # BASELINE configs[2]
    import bench
    import torch.distributed as dist
    from medicalseg_b200.models import VNet, VNetDeepSup, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    from medicalseg_b200.parallel import DistributedGradReducer

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg128 = bench.CONFIGS["vnet128_bf16"]
    if mri:
        model = (VNetDeepSup if mode == "deepsup" else VNet)(num_classes=20, compute_dtype="bf16", seed=0,
                     kernel_size=[[2, 2, 4], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
                     stride_size=[[2, 2, 1], [2, 2, 1], [2, 2, 2], [2, 2, 2]])
    else:
        model = VNet(num_classes=cfg128["classes"], compute_dtype=cdt, seed=0)
    model.train()
    losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    if mode == "deepsup":  # vnetdeepsup_mri_spine_seg_512_512_12_15k.yml: four MixedLoss objects x 0.25
        losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1]) for _ in range(4)],
                  "coef": [0.25] * 4}
    reducer = DistributedGradReducer(model.store.grad, bucket_mb=float(os.environ.get("MSB_BUCKET_MB", "32"))).attach(model)
    opt = Momentum(PolynomialDecay(0.001, 15000), model.parameters(), 0.9, 1e-4, grad_scale=reducer.grad_scale)
    img, lab = bench.synthetic_gpu_batch(cfg128, device, seed=rank)
    if mri:
        img = torch.rand(2, 1, 512, 512, 12, device=device)
        lab = torch.randint(0, 20, (2, 512, 512, 12), device=device, dtype=torch.int32)

    def step():
        logits_list = model(img)
        loss_list, dice = L.loss_computation(logits_list, lab, losses)
        loss = sum(loss_list)
        loss.backward()
        reducer.wait()
        opt.step()
        opt._learning_rate.step()
        model.clear_gradients()

    if evaluate:
        model.eval()
        train_step = step

        def step():
            with torch.no_grad():
                model.predict_with_losses(img, lab, losses)

    from medicalseg_b200 import _lib
    if os.environ.get("MSB_NO_PDL"):
        _lib.call("msb_debug_set", 7, 1)  # serialised kernels: per-kernel durations are not inflated by PDL waits
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    t_cpu = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10
    print("host enqueue time %.3f ms/step (python + launches, no sync); wall %.3f ms/step" % (t_cpu * 1e3, t_all * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if world > 1:
        dist.barrier()
        if rank == 0:
            # NCCL report: per step (delimited by momentum_kernel) the all-reduce kernels and the optimizer's wait
            step_start = ks[0][0] if ks else 0
            last_compute_end = None
            print("NCCL report (rank 0, world %d, bucket %s MB): all-reduce kernels [start offset in step us, duration us]"
                  % (world, os.environ.get("MSB_BUCKET_MB", "32")))
            cur = []
            for s_, e_, name in ks:
                if "nccl" in name.lower():
                    cur.append((round(s_ - step_start, 1), round(e_ - s_, 1)))
                elif "momentum_kernel" in name:
                    print("  step: span %.1f us, all-reduces %s, total NCCL kernel time %.1f us, optimizer started %.1f us "
                          "after the last backward kernel ended" % (e_ - step_start, cur, sum(d for _, d in cur),
                                                                    s_ - (last_compute_end or s_)))
                    cur, step_start = [], e_
                else:
                    last_compute_end = e_
        if rank != 0:
            dist.destroy_process_group()
            return
    if not ks:
        print("no CUDA events captured (CUPTI unavailable?)")
        return
    agg = collections.OrderedDict()
    busy = 0.0
    gaps = 0.0
    biggest = []
    last_end = ks[0][0]
    for s, e, name in ks:
        name = re.sub(r"\(.*$", "", name)
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += (e - s)
        busy += (e - s)
        if s > last_end:
            gaps += s - last_end
            biggest.append((s - last_end, name))
        last_end = max(last_end, e)
    span = ks[-1][1] - ks[0][0]
    print("traced %d steps: span %.3f ms/step, kernel time %.3f ms/step, idle gaps %.3f ms/step, %d launches/step" %
          (steps, span / steps / 1e3, busy / steps / 1e3, gaps / steps / 1e3, len(ks) // steps))
    print("%-92s %6s %9s %6s" % ("kernel", "n/step", "ms/step", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %6.1f %9.3f %5.1f%%" % (name[:92], n / steps, t / steps / 1e3, 100 * t / busy))
    biggest.sort(reverse=True)
    print("largest gaps (us, kernel that followed):", [(round(g, 1), n[:40]) for g, n in biggest[:8]])


if __name__ == "__main__":
    main()
