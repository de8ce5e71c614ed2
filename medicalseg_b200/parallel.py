"""Data-parallel gradient exchange (reference: core/train.py:81-88 — fleet.distributed_model / DataParallel).

One process per GPU; whole volumes are sharded across ranks; the ONLY collective on the training path is the
gradient all-reduce (NCCL over NVLink 5 / NVSwitch).  The model keeps all gradients in one flat f32 buffer laid
out in forward order, so backward completes it from the END towards the front: every transition block that
finishes its backward fires `grad_ready_hook(lo, hi)`, and we launch the all-reduce of that contiguous slice on a
side stream right away, overlapping it with the rest of backward.  Small slices are coalesced into buckets of at
least `bucket_mb` (default 8 MB; measured on 2 x B200: 12.30 ms/step with 32 MB buckets - the last ~30 MB were reduced
after backward had finished - vs 12.02 ms with 8 MB: 5 all-reduces, the last one 4.6 MB / 41 us).
Averaging (1/world) is folded into the optimizer kernel (Momentum.grad_scale) — no extra pass over the grads.

Two transports: `backend="direct"` (default on CUDA) calls ncclAllReduce on our own communicator (nccl.py) on the
reducer's side stream - plain stream-ordered work, so the whole data-parallel step can be captured into ONE CUDA graph
exactly like the single-GPU step; `backend="torch"` goes through torch.distributed (gloo in the CPU tests).

BatchNorm statistics: `VNet(sync_bn=True)` shares them across ranks like the reference's SyncBatchNorm conversion
(cvlibs/config.py:322; `train.py` enables it by default at world > 1); `sync_bn=False` keeps them per rank
(north_star: "allreduce for the gradient step only") - see DESIGN.md §6.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class BucketPlanner:
    """Pure host logic (testable on CPU/gloo): merges (lo, hi) ranges arriving in backward order into buckets
    of at least `min_elems` elements; ranges must tile the flat buffer from the end to the front."""

    def __init__(self, total: int, min_elems: int):
        self.total, self.min_elems = total, min_elems
        self.reset()

    def reset(self):
        self.pending_lo = self.pending_hi = self.total

    def add(self, lo: int, hi: int) -> Optional[Tuple[int, int]]:
        if hi != self.pending_lo:
            raise ValueError("gradient ranges must arrive contiguously from the end: got [%d,%d) after %d"
                             % (lo, hi, self.pending_lo))
        self.pending_lo = lo
        if self.pending_hi - self.pending_lo >= self.min_elems or lo == 0:
            out = (self.pending_lo, self.pending_hi)
            self.pending_hi = self.pending_lo
            return out
        return None

    def flush(self) -> Optional[Tuple[int, int]]:
        if self.pending_hi > self.pending_lo:
            out = (self.pending_lo, self.pending_hi)
            self.pending_hi = self.pending_lo
            return out
        return None


class DistributedGradReducer:
    """Attach to a model: reducer = DistributedGradReducer(model); after loss.backward() call reducer.wait()
    before optimizer.step().  Works with any torch.distributed backend (nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, flat_grad: torch.Tensor, bucket_mb: float = 8.0, group=None, backend: Optional[str] = None):
        self.flat_grad = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.planner = BucketPlanner(flat_grad.numel(), int(bucket_mb * (1 << 20) / 4))
        self.handles: List = []
        self.comm_stream = torch.cuda.Stream() if flat_grad.is_cuda else None
        self.launched: List[Tuple[int, int]] = []
        if backend is None:
            backend = "direct" if (flat_grad.is_cuda and group is None) else "torch"
        self.backend = backend
        self.comm = self.stat_comm = None
        if self.world > 1 and backend == "direct":
            from .nccl import Communicator
            self.comm = Communicator(flat_grad.device)

    @property
    def capturable(self) -> bool:
        """True when the all-reduces are plain stream-ordered NCCL calls that a CUDA-graph capture can record"""
        return self.world == 1 or self.comm is not None

    def all_reduce_(self, t: torch.Tensor):
        """small in-place sum on the CURRENT stream (SyncBatchNorm statistics); same transport as the gradients"""
        if self.world == 1:
            return
        if self.stat_comm is not None:
            self.stat_comm.all_reduce_(t)
        else:
            dist.all_reduce(t, group=self.group)

    def attach(self, model):
        model.grad_ready_hook = self.on_ready
        model.stat_all_reduce = self.all_reduce_
        if self.world > 1 and self.flat_grad.is_cuda and os.environ.get("MSB_TILE_SCHEDULER", "0") == "1":
            # opt-in experiment (needs a library built with MSB_DYNAMIC_TILES=1, raises otherwise): persistent conv grids
            # fetch tiles from an atomic counter (umma.cuh, namespace sched).  MEASURED on 2 x B200
            # (profiles/r2t_bench_n2*.log): 12.43 ms per step with it vs 12.41 ms without on the same box - the cost of
            # NCCL's CTAs is not a late-CTA tail effect, so static assignment is what ships
            from . import _lib
            _lib.call("msb_set_tile_scheduler", 1)
        if self.comm is not None and self.stat_comm is None and getattr(model, "sync_bn", False):
            # SyncBatchNorm sums run on the COMPUTE stream while gradient buckets are in flight on the side stream: one
            # NCCL communicator must not be used from two streams concurrently, so the statistics get their own
            from .nccl import Communicator
            self.stat_comm = Communicator(self.flat_grad.device)
        return self

    def _launch(self, lo: int, hi: int):
        self.launched.append((lo, hi))
        if self.world == 1:
            return
        view = self.flat_grad[lo:hi]
        if self.comm is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())  # fork (an event edge inside a capture)
            self.comm.all_reduce_(view, self.comm_stream)
        elif self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def on_ready(self, lo: int, hi: int):
        r = self.planner.add(lo, hi)
        if r is not None:
            self._launch(*r)

    def wait(self):
        r = self.planner.flush()
        if r is not None:
            self._launch(*r)
        for h in self.handles:
            h.wait()
        if self.comm_stream is not None and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.handles = []
        self.launched = []
        self.planner.reset()

    @property
    def grad_scale(self) -> float:
        """fold the 1/world averaging into the optimizer (paddle DataParallel averages gradients)"""
        return 1.0 / self.world
