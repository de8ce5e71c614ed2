"""GPU preprocessing with the reference signatures (tools/preprocess_utils/values.py:37-87, geometry.py:31-69).

Inputs may be numpy arrays (copied host->device, result returned as numpy — what `Prep.load_save` expects,
tools/prepare.py:236-249) or CUDA torch tensors (result stays on the device).  There is no CPU compute path."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("medicalseg_b200.preprocess needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(image, dtype):
    if torch.is_tensor(image):
        if not image.is_cuda:
            image = image.to(_device())
        return image.to(dtype).contiguous(), True
    arr = np.ascontiguousarray(np.asarray(image))
    return torch.from_numpy(arr).to(_device(), non_blocking=False).to(dtype).contiguous(), False


def _ret(t, was_tensor):
    return t if was_tensor else t.cpu().numpy()


def HUnorm(image, HU_min=-1200, HU_max=600, HU_nan=-2000):
    x, was = _to_dev(image, torch.float32)
    out = torch.empty_like(x)
    ops.hunorm(x, out, HU_min, HU_max, HU_nan)
    return _ret(out, was)


def normalize(image, min_val=None, max_val=None):
    x, was = _to_dev(image, torch.float32)
    out = torch.empty_like(x)
    if min_val is None and max_val is None:
        mm = torch.empty(2, dtype=torch.float32, device=x.device)
        ops.minmax(x, mm)
        ops.normalize(x, out, 0.0, 1.0, mm)
    else:
        ops.normalize(x, out, float(min_val), float(max_val), None)
    return _ret(out, was)


def label_remap(label, map_dict=None):
    x, was = _to_dev(label, torch.int32)
    x = x.clone()
    keys, vals = list(map_dict.keys()), list(map_dict.values())
    ops.label_remap(x, [int(k) for k in keys], [int(v) for v in vals])
    if not was:
        return x.cpu().numpy().astype(np.asarray(label).dtype)
    return x


def resample(image, spacing=None, new_spacing=(1.0, 1.0, 1.0), new_shape=None, order=1, pre_op=None):
    """geometry.py:31-69 -> (image_new, new_spacing).  `pre_op` (extension): ('hunorm', HU_min, HU_max, HU_nan) or
    ('normalize', lo, hi) fuses the value transform into the same gather pass (one read of the source)."""
    is_int = (torch.is_tensor(image) and not image.is_floating_point()) or \
             (not torch.is_tensor(image) and np.issubdtype(np.asarray(image).dtype, np.integer))
    shape = tuple(image.shape)
    if new_shape is None:
        sp = np.array([spacing[0], spacing[1], spacing[2]])
        new_shape = np.round(np.array(shape) * sp / np.asarray(new_spacing))
    else:
        new_shape = np.array(new_shape)
        if spacing is not None and len(spacing) == 4:
            spacing = spacing[1:]
        new_spacing = tuple((np.array(shape) / new_shape) * np.asarray(spacing)) if spacing is not None else None
    resize_factor = new_shape / np.array(shape)
    out_shape = [int(round(s * f)) for s, f in zip(shape, resize_factor)]
    if is_int:
        if order != 0:
            raise NotImplementedError("integer volumes are resampled with order=0 only (labels)")
        x, was = _to_dev(image, torch.int32)
        out = torch.empty(out_shape, dtype=torch.int32, device=x.device)
        ops.resample_i32(x, out)
        if not was:
            return out.cpu().numpy().astype(np.asarray(image).dtype), new_spacing
        return out, new_spacing
    x, was = _to_dev(image, torch.float32)
    out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    if pre_op is None:
        ops.resample_f32(x, out, order)
    elif pre_op[0] == "hunorm":
        ops.resample_f32(x, out, order, 1, *[float(v) for v in pre_op[1:4]])
    elif pre_op[0] == "normalize":
        ops.resample_f32(x, out, order, 2, float(pre_op[1]), float(pre_op[2]), 0.0)
    else:
        raise ValueError("unknown pre_op %r" % (pre_op,))
    return _ret(out, was), new_spacing


class ScanPipeline:
    """Host-resident scans -> preprocessed volumes with the copies hidden (BASELINE configs[4]: what `Prep.load_save`,
    tools/prepare.py:200-259, does per scan: images -> HUnorm + resample order 1, labels -> resample order 0).

    A 512^3 f32 scan + int32 label is 1.07 GB over PCIe against ~0.13 ms of kernels, so the pipeline is copy-bound; it
    keeps TWO device staging sets and runs three streams: H2D of scan i+1 (copy stream) overlaps the kernels of scan i
    (compute stream) and the D2H of its 128^3 results into pinned host buffers.  `run()` yields (image, label) pinned
    CPU tensors in order; a yielded pair stays valid until two more scans have been yielded.  Inputs should be pinned
    CPU tensors (a loader reading straight into pinned memory) - pageable inputs work but the copy then blocks."""

    def __init__(self, new_shape=(128, 128, 128), pre_op=("hunorm", -1200, 600, -2000), device=None, depth=2):
        self.device = torch.device(device) if device is not None else _device()
        self.new_shape, self.pre_op, self.depth = [int(v) for v in new_shape], pre_op, int(depth)
        self.h2d = torch.cuda.Stream(device=self.device)
        self.compute = torch.cuda.Stream(device=self.device)
        self.slots = [dict(img=None, lab=None, o_img=None, o_lab=None, h_img=None, h_lab=None, copied=None, done=None)
                      for _ in range(self.depth)]

    def _stage(self, slot, image, label):
        s = self.slots[slot]
        if s["done"] is not None:
            self.h2d.wait_event(s["done"])  # the kernels that read this staging set have finished
        if s["img"] is None or s["img"].shape != image.shape:
            s["img"] = torch.empty(image.shape, dtype=torch.float32, device=self.device)
            s["o_img"] = torch.empty(self.new_shape, dtype=torch.float32, device=self.device)
            s["h_img"] = torch.empty(self.new_shape, dtype=torch.float32, pin_memory=True)
        if label is not None and (s["lab"] is None or s["lab"].shape != label.shape):
            s["lab"] = torch.empty(label.shape, dtype=torch.int32, device=self.device)
            s["o_lab"] = torch.empty(self.new_shape, dtype=torch.int32, device=self.device)
            s["h_lab"] = torch.empty(self.new_shape, dtype=torch.int32, pin_memory=True)
        with torch.cuda.stream(self.h2d):
            s["img"].copy_(image, non_blocking=True)
            if label is not None:
                s["lab"].copy_(label, non_blocking=True)
            s["copied"] = torch.cuda.Event()
            s["copied"].record(self.h2d)
        s["has_label"] = label is not None

    def _process(self, slot):
        s = self.slots[slot]
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s["copied"])
            if self.pre_op is None:
                ops.resample_f32(s["img"], s["o_img"], 1)
            elif self.pre_op[0] == "hunorm":
                ops.resample_f32(s["img"], s["o_img"], 1, 1, *[float(v) for v in self.pre_op[1:4]])
            else:
                ops.resample_f32(s["img"], s["o_img"], 1, 2, float(self.pre_op[1]), float(self.pre_op[2]), 0.0)
            s["h_img"].copy_(s["o_img"], non_blocking=True)
            if s["has_label"]:
                ops.resample_i32(s["lab"], s["o_lab"])
                s["h_lab"].copy_(s["o_lab"], non_blocking=True)
            s["done"] = torch.cuda.Event()
            s["done"].record(self.compute)

    def run(self, scans):
        """scans: iterable of (image f32 [D,H,W] CPU tensor, label i32 [D,H,W] CPU tensor or None)"""
        with torch.cuda.device(self.device):
            it = iter(scans)
            pending = None
            k = 0
            for image, label in it:
                slot = k % self.depth
                self._stage(slot, image, label)       # H2D of scan k (overlaps the kernels / D2H of scan k-1)
                if pending is not None:
                    yield self._finish(pending)
                self._process(slot)
                pending = slot
                k += 1
            if pending is not None:
                yield self._finish(pending)

    def _finish(self, slot):
        s = self.slots[slot]
        s["done"].synchronize()
        return s["h_img"], (s["h_lab"] if s["has_label"] else None)
