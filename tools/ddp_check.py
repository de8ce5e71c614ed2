"""Data-parallel equivalence checks on >= 2 GPUs (reference: core/train.py:81-88 DataParallel + cvlibs/config.py:322
SyncBatchNorm).  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1
tools/ddp_check.py   (also reachable as `bench.py --gpus 2 --check`, which prints `ddp_grad_rel_err`).

 1. sync-BN + bucketed NCCL all-reduce: W ranks x batch 2 (statistics shared, gradients summed by the reducer's buckets
    on its own NCCL communicator) reproduce ONE process with batch 2W given the same d(loss)/d(logits): logits of the
    local half, summed flat gradient, running statistics.  f32 parity engine: <= 1e-4 / 1e-3 / 1e-5; bf16 engine:
    3e-2 / 0.15 / 1e-2 (rounding noise of the bf16 activations differs with the tile -> CTA order).
 2. the CAPTURED data-parallel step (all-reduces inside the CUDA graph) == the eager data-parallel step: 2 eager
    warm-up + 3 replays vs 5 eager steps, same batch and masks; and all ranks hold bit-identical parameters afterwards.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

SITES = (("down_tr128", 128), ("down_tr256", 256), ("up_tr256.x", 256), ("up_tr256.skip", 128), ("up_tr128.x", 256),
         ("up_tr128.skip", 64))


def _masks(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: (torch.rand(batch, c, generator=g) >= 0.5).float() * 2.0 for k, c in SITES}


def sync_bn_and_reducer(dtype, dev, rank, world, shape=(16, 16, 16)):
    from medicalseg_b200.models import VNet
    from medicalseg_b200.parallel import DistributedGradReducer
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2 * world, 1, *shape, generator=g)
    dlog = torch.randn(2 * world, 2, *shape, generator=g)
    masks = _masks(2 * world, 1)
    lo, hi = 2 * rank, 2 * rank + 2
    m = VNet(num_classes=2, compute_dtype=dtype, seed=0, sync_bn=True)
    red = DistributedGradReducer(m.store.grad, bucket_mb=8.0).attach(m)
    m.train()
    m.set_dropout_masks({k: v[lo:hi] for k, v in masks.items()})
    logits = m._forward(img[lo:hi].to(dev), record=True)
    m._backward(dlog[lo:hi].to(dev))
    nb = len(red.launched) + (1 if red.planner.pending_hi > red.planner.pending_lo else 0)
    red.wait()
    ref = VNet(num_classes=2, compute_dtype=dtype, seed=0, sync_bn=False)
    ref.train()
    ref.set_dropout_masks(masks)
    rlogits = ref._forward(img.to(dev), record=True)
    ref._backward(dlog.to(dev))
    tol_l, tol_g, tol_b = (1e-4, 1e-3, 1e-5) if dtype == "f32" else (3e-2, 0.15, 1e-2)
    e_log = float((logits - rlogits[lo:hi]).abs().max() / rlogits.abs().max())
    e_grad = float((m.store.grad - ref.store.grad).norm() / ref.store.grad.norm())
    e_buf = float((m.store.buffers - ref.store.buffers).abs().max())
    ok = e_log <= tol_l and e_grad <= tol_g and e_buf <= tol_b
    print("rank %d %s: %d all-reduce buckets (backend %s); logits rel err %.3g, ddp_grad_rel_err %.3g, running-stat max "
          "diff %.3g -> %s" % (rank, dtype, nb, red.backend, e_log, e_grad, e_buf, "OK" if ok else "MISMATCH"), flush=True)
    return ok, e_grad


def graph_vs_eager(dev, rank, world, shape=(32, 32, 32)):
    from medicalseg_b200.graph import GraphedTrainStep
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    from medicalseg_b200.parallel import DistributedGradReducer
    g = torch.Generator().manual_seed(100 + rank)  # every rank trains on its own volumes
    img = torch.rand(2, 1, *shape, generator=g).to(dev)
    lab = (torch.rand(2, *shape, generator=g) > 0.5).to(torch.int32).to(dev)
    masks = _masks(2, 7 + rank)

    def make():
        m = VNet(num_classes=2, compute_dtype="bf16", seed=0, sync_bn=True)
        m.train()
        m.set_dropout_masks(masks, persistent=True)
        losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
        red = DistributedGradReducer(m.store.grad, bucket_mb=8.0).attach(m)
        opt = Momentum(PolynomialDecay(0.01, 100), m.parameters(), 0.9, 1e-4, grad_scale=red.grad_scale)
        return m, losses, opt, red

    m1, l1, o1, r1 = make()
    p0 = m1.store.flat.clone()
    eager = []
    for _ in range(5):
        ll, _ = L.loss_computation(m1(img), lab, l1)
        loss = sum(ll)
        loss.backward()
        r1.wait()
        o1.step(); o1._learning_rate.step(); m1.clear_gradients()
        eager.append(float(loss))
    m2, l2, o2, r2 = make()
    step = GraphedTrainStep(m2, l2, o2, warmup=2, reducer=r2)
    graph = []
    for _ in range(3):
        loss, _ = step(img, lab)
        graph.append(float(loss))
    torch.cuda.synchronize()
    moved = float((m1.store.flat - p0).abs().max())
    e_par = float((m1.store.flat - m2.store.flat).abs().max()) / max(moved, 1e-30)
    e_loss = max(abs(a - b) / abs(a) for a, b in zip(eager[2:], graph))
    # replicas stay bit-identical: the all-reduced gradient is the same bits on every rank
    mx, mn = m2.store.flat.clone(), m2.store.flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    spread = float((mx - mn).abs().max())
    ok = step.captured and e_par <= 5e-3 and e_loss <= 5e-3 and spread == 0.0
    print("rank %d graph-vs-eager (world %d): loss rel err %.3g, param err / largest update %.3g, replica spread %.3g "
          "-> %s" % (rank, world, e_loss, e_par, spread, "OK" if ok else "MISMATCH"), flush=True)
    return ok


def run_checks(dev, rank, world, with_bf16=True):
    """returns (ok, ddp_grad_rel_err of the f32 engine); the caller owns the process group"""
    ok, e_grad = sync_bn_and_reducer("f32", dev, rank, world)
    if with_bf16:
        ok = sync_bn_and_reducer("bf16", dev, rank, world)[0] and ok
    ok = graph_vs_eager(dev, rank, world) and ok
    t = torch.tensor([0.0 if ok else 1.0], device=dev)
    dist.all_reduce(t)
    return float(t.item()) == 0.0, e_grad


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    ok, e = run_checks(dev, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("DDP_CHECK_OK ddp_grad_rel_err=%.3g" % e)


if __name__ == "__main__":
    main()
