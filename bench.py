#!/usr/bin/env python
"""bench.py — VNet 128^3 train-step volumes/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5                # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # reference arm: CPU restatement on host cores

One JSON line on stdout (rank 0).  A "step" = forward + CE/Dice loss + backward + gradient all-reduce (N>1) +
Momentum update of one batch of 2 synthetic 128^3 volumes per GPU (config 1 of BASELINE.json: "VNet 128^3 2-class
bf16 train step, batch 2, 1xB200"); `value` times it with inputs resident in HBM, `e2e` through the public API with
pinned HOST inputs (H2D inside the timed region) and the loss read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "VNet 128^3 train-step volumes/sec"
UNIT = "volumes/s"
SHAPE = (128, 128, 128)
BATCH = 2
NUM_CLASSES = 2
# conv/convT MACs x2, fwd + dgrad + wgrad, 128^3, C=2 (per volume)
GFLOP_PER_VOLUME = 4380.9
DOMINANT_GFLOP_PER_VOLUME = 536.87  # up_tr32.ops[0].conv1, 32->32 5^3 @128^3, one pass


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        """windows: [(t0, t1), ...] wall-clock intervals during which the GPU was running timed steps"""
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for ts, r in self.rows if len(r) >= 8 and any(a <= ts <= b for a, b in windows)]
        sm = sorted(int(float(r[1])) for r in rows if r[1].replace(".", "").isdigit())
        mx = max([int(float(r[2])) for r in rows if r[2].replace(".", "").isdigit()] or [0])
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "sampled_over": "device-resident and e2e timed regions (nvidia-smi -lms 100)"}


def synthetic_gpu_batch(device, seed):
    """same recipe as oracle.vnet_oracle.synthetic_batch (SURVEY.md §8d cfg 2), generated on the device"""
    g = torch.Generator(device=device).manual_seed(seed)
    img = torch.rand(BATCH, 1, *SHAPE, generator=g, device=device)
    img = img / img.amax(dim=(1, 2, 3, 4), keepdim=True)
    noise = torch.rand(BATCH, 1, *SHAPE, generator=g, device=device)
    sm = noise
    for _ in range(2):
        sm = torch.nn.functional.avg_pool3d(torch.nn.functional.pad(sm, (2,) * 6, mode="replicate"), 5, stride=1)
    thr = sm.flatten(1).median(dim=1).values.view(-1, 1, 1, 1)
    lab = (sm[:, 0] > thr).to(torch.int32)
    return img.contiguous(), lab.contiguous()


def run_ours(args):
    import torch.distributed as dist
    from medicalseg_b200 import _lib
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    from medicalseg_b200.parallel import DistributedGradReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    model = VNet(num_classes=NUM_CLASSES, compute_dtype="bf16", seed=0)
    model.train()
    losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    reducer = DistributedGradReducer(model.store.grad).attach(model)
    opt = Momentum(PolynomialDecay(0.001, 15000), model.parameters(), 0.9, 1e-4, grad_scale=reducer.grad_scale)
    img, lab = synthetic_gpu_batch(device, seed=rank)
    orig_call = _lib.call
    # world == 1: the whole step is ONE CUDA graph (medicalseg_b200.graph.GraphedTrainStep, the `to_static_training`
    # hook of core.train); world > 1 keeps the eager path (bucketed NCCL all-reduce overlapping backward)
    use_graph = world == 1 and not args.no_graph
    gstep = None
    if use_graph:
        from medicalseg_b200.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, losses, opt, reducer=reducer)

    def eager_step(images, labels):
        logits_list = model(images)
        loss_list, dice = L.loss_computation(logits_list, labels, losses)  # dice: lazy D2H (read on first use)
        loss = sum(loss_list)
        loss.backward()
        reducer.wait()
        opt.step()
        opt._learning_rate.step()
        model.clear_gradients()
        return loss, dice

    def step(images, labels):
        return gstep(images, labels) if use_graph else eager_step(images, labels)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that nvidia-smi is already streaming in the timed regions
    # ---- count this library's C-ABI calls of ONE step (every call launches >= 1 kernel of libmedseg_b200.so) ----------
    import medicalseg_b200.ops as ops_mod
    counter = {"n": 0}

    def counting_call(name, *a):
        counter["n"] += 1
        return orig_call(name, *a)

    eager_step(img, lab)
    ops_mod.call = counting_call
    eager_step(img, lab)
    ops_mod.call = orig_call
    calls_per_step = counter["n"]
    for _ in range(args.warmup):
        step(img, lab)
    barrier()

    # ---- device-resident timing (value) -------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    windows = []
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step(img, lab)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- end to end through the public API with HOST inputs (e2e) -----------------------------------------
    h_img = img.cpu().pin_memory()
    h_lab = lab.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 30))
    last = None
    if use_graph:
        # input pipelining: the H2D copy of batch i+1 (pinned host memory -> staging buffers, copy stream) is issued
        # right after the replay of step i was launched, so it overlaps that step; every step still copies its inputs
        # H2D and reads its loss D2H inside the timed region
        def e2e_loop(k):
            nonlocal last
            gstep.prefetch(h_img, h_lab)
            for _ in range(k):
                loss, dice = gstep()
                gstep.prefetch(h_img, h_lab)
                last = float(loss.item())  # D2H read of the step's result
    else:
        # eager path (world > 1): the same input pipelining through medicalseg_b200.utils.DevicePrefetcher
        from medicalseg_b200.utils import DevicePrefetcher
        pre = DevicePrefetcher(device)

        def e2e_loop(k):
            nonlocal last
            pre.stage(h_img, h_lab)
            for _ in range(k):
                d_img, d_lab = pre.get()
                pre.stage(h_img, h_lab)  # H2D of the next batch overlaps this step
                loss, dice = step(d_img, d_lab)
                last = float(loss.item())  # D2H read of the step's result
            pre.get()
    e2e_loop(3)  # untimed warm-up of the end-to-end path (staging buffers, copy stream, pinned result buffers)
    barrier()
    w0 = time.time()
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    clocks = sampler.stop(windows) if rank == 0 else None
    t2 = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / e2e_steps

    out = None
    if rank == 0:
        burst, sustained, hbm, src = peaks()
        # ---- roofline of the dominant kernel: the 32->32 5x5x5 conv at 128^3 (up_tr32.ops[0].conv1), timed alone
        from medicalseg_b200 import ops
        from medicalseg_b200.ops import B8
        lu = model.up_tr32.ops[0]
        xb = B8(BATCH, 32, SHAPE, torch.bfloat16, device=device)
        xb.buf.normal_()
        yb = B8(BATCH, 32, SHAPE, torch.bfloat16, device=device)
        lu.k5._pack(32, 32)
        sums = torch.zeros(64, dtype=torch.float64, device=device)
        for _ in range(3):
            ops.k5_fwd(xb, lu.k5.packed_f, model.store.view(lu.conv1.bias), 32, yb, False, None, 1, sums)
        torch.cuda.synchronize()
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.k5_fwd(xb, lu.k5.packed_f, model.store.view(lu.conv1.bias), 32, yb, False, None, 1, sums)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        achieved = BATCH * DOMINANT_GFLOP_PER_VOLUME / k_ms  # GFLOP / ms = TFLOP/s
        del xb, yb
        vols = world * BATCH * args.steps / (ms / 1e3)
        out = {
            "metric": METRIC, "value": round(vols, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "VNet(num_classes=2) 128x128x128 bf16 train step (fwd + CE/Dice + bwd + Momentum), "
                                   "batch 2 per GPU, BASELINE.json configs[1]",
                       "global_batch": world * BATCH, "parallelism": "dp%d" % world,
                       "l2_policy": "per-step working set (~3 GB activations) is far larger than the 126 MB L2",
                       "bn_statistics": "per-rank batch statistics (no collective in fwd/bwd)"},
            "e2e": {"value": round(world * BATCH / (e2e_ms / 1e3), 3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h_img.numel() * 4 + h_lab.numel() * 4),
                    "d2h_bytes_per_step": int(4 * (1 + 2 + NUM_CLASSES)), "ms_per_step": round(e2e_ms, 3),
                    "last_loss": last},
            # C-ABI calls of one eager step x timed steps (a graph replay launches the same kernels; several calls
            # launch 2-3 kernels, so this is a lower bound of the kernel count)
            "gpu_launches": calls_per_step * args.steps,
            "launch_mode": "cuda-graph replay (1 cudaGraphLaunch/step)" if use_graph else "eager (python -> C ABI)",
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": burst, "unit": "TFLOP/s",
                         "frac": round(achieved / burst, 4),
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this shape
                         # (ncu --set full, profiles/r1s_ncu_full_fwd32_wgrad32_final.txt): 268.7 + 224.1 MB, vs 536.9 MB
                         # of algorithmic activation bytes (x + y, bf16) -> every byte moves once
                         "traffic": 492.9e6, "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "peak_source": src + " (burst, kernel timed alone)",
                         "kernel": "conv_k5_fwd_kernel<32,8,8> (up_tr32.ops[0].conv1 32->32 5x5x5 @128^3, batch 2)",
                         "kernel_ms": round(k_ms, 4),
                         "step_tflops": round(world * BATCH * args.steps * GFLOP_PER_VOLUME / ms, 2),
                         "step_frac_of_sustained": round(BATCH * args.steps * GFLOP_PER_VOLUME / ms / sustained, 4)},
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(sample_depth=args.cpu_depth)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def cpu_step_time(depth, steps, warmup, batch=1):
    """times the oracle (torch-CPU restatement of the reference; PaddlePaddle cannot be installed offline) on a
    [batch,1,depth,128,128] slab of the 128^3 workload; returns (seconds per step, cores)"""
    from oracle import vnet_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = vo.VNetOracle(num_classes=NUM_CLASSES)
    model.train()
    losses = vo.default_losses()
    opt = vo.Momentum(vo.PolynomialDecay(0.001, 15000), list(model.parameters()), 0.9, 1e-4)
    img, lab = vo.synthetic_batch(batch, (depth, SHAPE[1], SHAPE[2]), NUM_CLASSES, seed=0)
    times = []
    for i in range(warmup + steps):
        masks = vo.make_dropout_masks(batch, seed=0, step=i)
        t0 = time.perf_counter()
        vo.train_step(model, losses, opt, img, lab, masks)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def cpu_baseline(sample_depth=32):
    sec, cores = cpu_step_time(sample_depth, steps=1, warmup=1)
    frac = sample_depth / SHAPE[0]
    return {"value": round(frac / sec, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle (torch-CPU f32 restatement of the reference VNet; PaddlePaddle not installable offline): "
                      "1 timed train step (after 1 warm-up) on a batch-1 %dx128x128 slab = %.3f of a 128^3 volume, "
                      "%.2f s/step" % (sample_depth, frac, sec)}


def run_reference(args):
    """reference arm: the reference's CPU path (oracle port) on the box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    depth = args.cpu_depth
    sec, cores = cpu_step_time(depth, steps=args.steps, warmup=min(args.warmup, 1))
    frac = depth / SHAPE[0]
    val = round(frac / sec, 5)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "VNet(num_classes=2) 128x128x128 train step, reference CPU path (torch-CPU restatement; "
                               "PaddlePaddle not installable offline), bounded sample: batch-1 %dx128x128 slab per step"
                               % depth},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "batch-1 %dx128x128 slab (%.3f volume) per step, %d timed steps" % (depth, frac, args.steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph (N=1)")
    ap.add_argument("--cpu-depth", type=int, default=32, help="depth of the 128x128 slab the CPU arm processes per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
