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
