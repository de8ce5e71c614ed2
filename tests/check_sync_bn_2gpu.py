"""(test infrastructure: lives under tests/ because it uses the oracle's synthetic batch / dropout-mask helpers)
2-GPU check of VNet(sync_bn=True): two ranks with batch 2 each must reproduce ONE process with batch 4 - the
reference converts every BatchNorm to SyncBatchNorm when world > 1 (cvlibs/config.py:322).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/check_sync_bn_2gpu.py

Forward: logits of the local half == the matching half of the single-process logits.  Backward (given the same
d(loss)/d(logits)): the SUM over ranks of the flat gradient buffers == the single-process gradients, running statistics
identical.  f32 parity path: tolerance 1e-4 (logits, relative to max|logit|) / 1e-3 (gradients, relative to the norm);
measured on 2 x B200: logits identical, gradients 5e-8.  bf16 (optional argument): the rank whose volumes sit at batch
positions 0,1 in both runs reproduces bit-identical results, the other one differs by bf16 rounding noise (different
tile -> CTA order): logits 8e-3, gradients 6e-2 of the norm - tolerance 3e-2 / 0.15.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from medicalseg_b200.models import VNet
    from oracle import vnet_oracle as vo  # test infrastructure: deterministic synthetic batch + dropout masks

    dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
    shape = (16, 16, 16)
    img, _ = vo.synthetic_batch(2 * world, shape, 2, seed=0)
    masks = vo.make_dropout_masks(2 * world, seed=0)
    g = torch.Generator().manual_seed(7)
    dlog = torch.randn(2 * world, 2, *shape, generator=g)
    lo, hi = 2 * rank, 2 * rank + 2

    # ---- this rank: its half of the batch, statistics shared through NCCL
    m = VNet(num_classes=2, compute_dtype=dtype, seed=0, sync_bn=True)
    m.train()
    m.set_dropout_masks({k: v[lo:hi] for k, v in masks.items()})
    logits = m._forward(img[lo:hi].to(dev), record=True)
    m._backward(dlog[lo:hi].to(dev))
    gsum = m.store.grad.clone()
    dist.all_reduce(gsum)

    # ---- single process, whole batch (every rank computes it redundantly; no collectives: sync_bn off)
    ref = VNet(num_classes=2, compute_dtype=dtype, seed=0, sync_bn=False)
    ref.train()
    ref.set_dropout_masks(masks)
    rlogits = ref._forward(img.to(dev), record=True)
    ref._backward(dlog.to(dev))

    tol_l, tol_g = (1e-4, 1e-3) if dtype == "f32" else (3e-2, 0.15)
    e_log = float((logits - rlogits[lo:hi]).abs().max() / rlogits.abs().max())
    e_grad = float((gsum - ref.store.grad).norm() / ref.store.grad.norm())
    e_buf = float((m.store.buffers - ref.store.buffers).abs().max())
    ok = e_log <= tol_l and e_grad <= tol_g and e_buf <= (1e-5 if dtype == "f32" else 1e-2)
    print("rank %d %s: logits rel err %.3g, summed-gradient rel err %.3g, running-stat max diff %.3g -> %s" %
          (rank, dtype, e_log, e_grad, e_buf, "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
