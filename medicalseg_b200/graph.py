"""One CUDA graph for the whole train step (reference hook: `to_static_training`, core/train.py:45 / train.py:88 -
the reference offers Paddle's @to_static graph mode there; here the step is captured into a CUDA graph).

The eager step issues ~320 kernel launches from Python; at 13 ms per step that is fine while the host runs ahead, but
every host synchronisation (the reference reads the loss every iteration, core/train.py:158) lets the GPU drain and the
next step then starts at the pace of the Python launch loop.  Captured once, a step is ONE cudaGraphLaunch:

    step = GraphedTrainStep(model, losses, optimizer)      # captures on the first call (after eager warm-up steps)
    loss, per_channel_dice = step(images, labels)          # H2D/D2D into static buffers + graph replay

Everything that varies between steps lives in device memory: the batch (static input buffers), the learning rate
(optimizer.lr_dev, msb_momentum_step_lrdev), the Dropout3D masks (torch's graph-safe Philox generator).  Weight
re-packing for the tensor-core kernels is issued at the START of the captured step on a side stream (each conv waits
for its own layer), so it overlaps the first convolutions as in the eager path.

world_size > 1: the bucketed gradient all-reduces are captured too.  Capturing torch.distributed's ProcessGroupNCCL
work objects deadlocked on 2 x B200 in round 1 (its own streams / events / watchdog bookkeeping inside the capture), so
the reducer now owns a plain NCCL communicator (nccl.py) and every bucket is one `ncclAllReduce` on the reducer's side
stream, forked from / joined to the capturing stream by events - the data-parallel step is the same single graph launch
as the single-GPU step, with the collectives overlapping backward inside the graph.  A reducer on the
torch.distributed transport (gloo CPU tests) is not capturable and is rejected.
"""
from __future__ import annotations

import torch

from .models import losses as L


class GraphedTrainStep:
    def __init__(self, model, losses, optimizer, warmup: int = 3, reducer=None):
        self.model, self.losses, self.optimizer, self.warmup = model, losses, optimizer, warmup
        self.reducer = reducer
        self.graph = None
        self.s_img = self.s_lab = self.s_loss = self.s_dice = None
        self._copy_stream = self._staged = self._consumed = None
        self.g_img = self.g_lab = None

    # ---- one eager step on the static buffers (also the body that gets captured) --------------------------------
    def _body(self):
        m, opt = self.model, self.optimizer
        if m._defer_prepack:
            m.param_version += 1       # parameters were updated by the previous replay's optimizer step
            m.prepack_async()
        # The model's forward/backward are explicit kernel sequences (VNet._forward / _backward); only the loss goes
        # through autograd, on a leaf created INSIDE the capture: autograd's end-of-backward stream synchronisation
        # then only ever sees the capturing stream (a leaf first used on another stream - the model's autograd anchor
        # - would make the engine wait on uncaptured work: cudaErrorStreamCaptureIsolation).
        with torch.no_grad():
            logits = m._forward(self.s_img, record=True)
            extra = list(m._extra_logits) if getattr(m, "deep_supervision", False) else []
        # VNetDeepSup: [out, d1, d2, d3] (vnet_deepsup.py:256-275) - one leaf per output, the heads' gradients enter the
        # trunk's backward through the same hooks the eager autograd function uses
        leaves = [t.detach().requires_grad_(True) for t in [logits] + extra]
        loss_list, dice = L.loss_computation(leaves, self.s_lab, self.losses)
        loss = sum(loss_list)
        loss.backward()
        if extra:
            m._aux_backward_begin(tuple(l.grad for l in leaves[1:]))
        m._backward(leaves[0].grad)
        if self.reducer is not None:
            self.reducer.wait()  # joins the all-reduce side stream (world > 1) and resets the bucket planner
        opt.step()
        m.clear_gradients()
        if m._defer_prepack and m._side_stream is not None:
            torch.cuda.current_stream().wait_stream(m._side_stream)  # join the side stream inside the capture
        return loss.detach(), dice

    def _capture(self, images, labels):
        if self.reducer is not None and not getattr(self.reducer, "capturable", True):
            raise RuntimeError("GraphedTrainStep at world_size > 1 needs DistributedGradReducer(backend='direct') (plain "
                               "ncclAllReduce calls a capture can record); torch.distributed work objects are not "
                               "capturable (see the module docstring)")
        m, opt = self.model, self.optimizer
        dev = m.device
        self.s_img = torch.empty_like(images, device=dev)
        self.s_lab = torch.empty_like(labels, device=dev, dtype=torch.int32)
        self.s_img.copy_(images)
        self.s_lab.copy_(labels)
        opt.lr_dev = torch.full((1,), float(opt.get_lr()), dtype=torch.float32, device=dev)
        # eager warm-up on a side stream (torch.cuda.graph requirement): first-call caches (CE class weights, packed
        # operand buffers, workspaces, cudaFuncSetAttribute) are all populated before the capture.  These are REAL
        # steps of the same batch, so the caller counts them (see __call__).
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        out = None
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                out = self._body()
                self._advance_lr()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for pk in getattr(m, "_packers", []):
            pk.pack_event = None  # completed (synchronize above); a capturing stream must not wait on outside events
        m._defer_prepack = True
        self.graph = torch.cuda.CUDAGraph()
        # thread-local error mode: other threads (torch.distributed's NCCL watchdog, the clock sampler) may call CUDA
        # APIs while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            loss, dice = self._body()
        self.s_loss = loss
        self.s_dice = dice._dev if isinstance(dice, L.LazyHostArray) else None
        return out

    def _advance_lr(self):
        opt = self.optimizer
        if hasattr(opt._learning_rate, "step"):
            opt._learning_rate.step()
        opt.lr_dev.fill_(float(opt.get_lr()))

    @property
    def captured(self) -> bool:
        return self.graph is not None

    def prefetch(self, images, labels):
        """starts the host->device copy of the NEXT batch on a copy stream, so that it overlaps the step that is
        running; the next `step()` call without arguments consumes it.  (pinned host tensors make it asynchronous)"""
        if self.graph is None:
            raise RuntimeError("prefetch() needs a captured step: call step(images, labels) once first")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.model.device)
            self.g_img, self.g_lab = torch.empty_like(self.s_img), torch.empty_like(self.s_lab)
        st = self._copy_stream
        if self._consumed is not None:
            st.wait_event(self._consumed)  # the previously staged batch has been moved into the graph's inputs
        with torch.cuda.stream(st):
            self.g_img.copy_(images, non_blocking=True)
            self.g_lab.copy_(labels, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record(st)

    def __call__(self, images=None, labels=None):
        """runs ONE train step (forward, loss, backward, optimizer step, LR-schedule step, clear_gradients) and
        returns (loss tensor, per-class Dice LazyHostArray).  Without arguments it consumes the batch staged by
        prefetch().  The first call captures the graph; its warm-up steps are `self.warmup` extra REAL optimizer steps
        on that first batch (the LR schedule advances accordingly)."""
        if self.graph is None:
            self._capture(images, labels)
        elif images is None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged)
            self.s_img.copy_(self.g_img, non_blocking=True)
            self.s_lab.copy_(self.g_lab, non_blocking=True)
            self._consumed = torch.cuda.Event()
            self._consumed.record(cur)
        else:
            self.s_img.copy_(images, non_blocking=True)
            self.s_lab.copy_(labels, non_blocking=True)
        self.graph.replay()
        # the replay re-packed the tensor-core weight images BEFORE its optimizer step: tell the Python side that they
        # are one step old, so that an eager forward (evaluation, odd-shaped batch) re-packs lazily
        self.model.param_version += 1
        self._advance_lr()
        dice = L.LazyHostArray(self.s_dice) if self.s_dice is not None else None
        return self.s_loss, dice
