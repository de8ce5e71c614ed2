"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference preprocessing ops.

PINNED: tests/golden/make_golden.py runs the reference's own tools/preprocess_utils/values.py and
geometry.py (NumPy/SciPy branch) in the build container and stores their outputs in
tests/golden/preprocess_*.npz; tests/test_oracle_preprocess.py checks this restatement against them.

scipy.ndimage.zoom (SciPy 1.x, `grid_mode=False`, no prefilter for order<=1) is a third-party
dependency of the reference (requirements.txt: scipy, unpinned).  Its published algorithm, restated:
  out_shape_i = round(in_i * zoom_i);  src_i = dst_i * (in_i - 1) / (out_i - 1)   (0 when out_i == 1)
  order 0: value at floor(src + 0.5);  order 1: multilinear blend of floor(src) and floor(src)+1,
  both clamped to the edge (mode='nearest'); arithmetic in float64, result cast to the input dtype
  (rounded for integer dtypes).
"""
from __future__ import annotations

import numpy as np


def label_remap(label, map_dict=None):  # tools/preprocess_utils/values.py:37-51
    label = np.array(label, copy=True)
    for key, val in map_dict.items():  # sequential in-place, as the reference
        label[label == key] = val
    return label


def normalize(image, min_val=None, max_val=None):  # values.py:54-64
    image = np.asarray(image)
    if min_val is None and max_val is None:
        image = (image - image.min()) / (image.max() - image.min())
    else:
        image = (image - min_val) / (max_val - min_val)
    return np.clip(image, 0, 1)


def HUnorm(image, HU_min=-1200, HU_max=600, HU_nan=-2000):  # values.py:67-87
    image = np.array(image, copy=True)
    image = np.nan_to_num(image, copy=False, nan=HU_nan)
    image = (image - HU_min) / ((HU_max - HU_min) / 255)
    return np.clip(image, 0, 255)


def _axis_coords(n_in: int, n_out: int) -> np.ndarray:
    if n_out <= 1:
        return np.zeros(max(n_out, 0), dtype=np.float64)
    return np.arange(n_out, dtype=np.float64) * ((n_in - 1) / (n_out - 1))


def zoom_to_shape(image: np.ndarray, out_shape, order: int) -> np.ndarray:
    """scipy.ndimage.zoom(image, out/in, mode='nearest', order) for order in {0, 1}, 3-D input."""
    assert image.ndim == 3 and order in (0, 1)
    out = image.astype(np.float64)
    for ax in range(3):
        n_in, n_out = image.shape[ax], int(out_shape[ax])
        src = _axis_coords(n_in, n_out)
        if order == 0:
            idx = np.clip(np.floor(src + 0.5).astype(np.int64), 0, n_in - 1)
            out = np.take(out, idx, axis=ax)
        else:
            i0 = np.floor(src).astype(np.int64)
            t = src - i0
            i1 = np.clip(i0 + 1, 0, n_in - 1)
            i0 = np.clip(i0, 0, n_in - 1)
            shp = [1, 1, 1]
            shp[ax] = n_out
            t = t.reshape(shp)
            out = np.take(out, i0, axis=ax) * (1.0 - t) + np.take(out, i1, axis=ax) * t
    if np.issubdtype(image.dtype, np.integer):
        out = np.rint(out)
    return out.astype(image.dtype)


def resample(image, spacing=None, new_spacing=(1.0, 1.0, 1.0), new_shape=None, order=1):
    """tools/preprocess_utils/geometry.py:31-69 -> (image_new, new_spacing)."""
    image = np.asarray(image)
    if new_shape is None:
        spacing = np.array([spacing[0], spacing[1], spacing[2]])
        new_shape = np.round(image.shape * spacing / np.asarray(new_spacing))
    else:
        new_shape = np.array(new_shape)
        if spacing is not None and len(spacing) == 4:
            spacing = spacing[1:]
        new_spacing = tuple((image.shape / new_shape) * spacing) if spacing is not None else None
    resize_factor = new_shape / np.array(image.shape)
    out_shape = [int(round(s * f)) for s, f in zip(image.shape, resize_factor)]
    return zoom_to_shape(image, out_shape, order), new_spacing
