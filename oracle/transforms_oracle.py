"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference's training augmentations
(medicalseg/transforms/functional.py:25-110, transform.py:27-339) for the transforms the shipped configs use.

PINNED: tests/golden/make_transforms_golden.py imports the reference's OWN transform.py / functional.py in the build
container (SimpleITK and the component manager stubbed - neither is touched by these classes), runs them with seeded
`random` / `numpy.random` and stores inputs + outputs in tests/golden/transforms_ref.npz;
tests/test_oracle_transforms.py checks this restatement against them.

scipy.ndimage.rotate / zoom are third-party dependencies of the reference (requirements.txt: scipy, unpinned).  The
published algorithm of rotate(reshape=False, mode='constant', order<=1), restated (no SciPy call below except the
degree-exact cosdg / sindg it uses itself):
  c, s = cosdg(angle), sindg(angle);  M = [[c, s], [-s, c]];  off = (n-1)/2 - M @ (n-1)/2   (plane extents n)
  input coordinate of output voxel o:  x = (o_a*M[.,0] + o_b*M[.,1]) + off      (f64, this order)
  x outside 0 <= x <= n-1 on either plane axis -> cval; order 0 -> voxel floor(x+0.5); order 1 -> bilinear blend of
  floor(x), floor(x)+1 in f64; integer arrays are rounded (floor(v+0.5)) back to their dtype.
"""
from __future__ import annotations

import numbers
import random

import numpy as np

from . import preprocess_oracle as po


def _cos_sin_degrees(angle):
    from scipy import special  # the same degree-exact functions scipy.ndimage.rotate calls
    return float(special.cosdg(angle)), float(special.sindg(angle))


def rotate_3d(img, r_plane, angle, order=1, cval=0):  # functional.py:91-100
    img = np.asarray(img)
    a, b = sorted(int(x) for x in r_plane)
    c, s = _cos_sin_degrees(angle)
    m = np.array([[c, s], [-s, c]])
    na, nb = img.shape[a], img.shape[b]
    center = (np.array([na, nb], dtype=np.float64) - 1) / 2
    off = center - m @ center
    grid = np.indices(img.shape)
    oa, ob = grid[a].astype(np.float64), grid[b].astype(np.float64)
    xa = (oa * m[0, 0] + ob * m[0, 1]) + off[0]
    xb = (oa * m[1, 0] + ob * m[1, 1]) + off[1]
    inside = (xa >= 0) & (xa <= na - 1) & (xb >= 0) & (xb <= nb - 1)

    def fetch(ia, ib):
        sel = [grid[0], grid[1], grid[2]]
        sel[a], sel[b] = np.clip(ia, 0, na - 1), np.clip(ib, 0, nb - 1)
        return img[tuple(sel)].astype(np.float64)

    if order == 0:
        v = fetch(np.floor(xa + 0.5).astype(np.int64), np.floor(xb + 0.5).astype(np.int64))
    elif order == 1:
        fa, fb = np.floor(xa), np.floor(xb)
        ta, tb = xa - fa, xb - fb
        fa, fb = fa.astype(np.int64), fb.astype(np.int64)
        v = (1 - ta) * (1 - tb) * fetch(fa, fb)
        v = v + (1 - ta) * tb * fetch(fa, fb + 1)
        v = v + ta * (1 - tb) * fetch(fa + 1, fb)
        v = v + ta * tb * fetch(fa + 1, fb + 1)
    else:
        raise NotImplementedError("orders above 1 need SciPy's spline prefilter")
    v = np.where(inside, v, float(cval))
    if np.issubdtype(img.dtype, np.integer):
        v = np.floor(v + 0.5)
    return v.astype(img.dtype)


def flip_3d(img, axis):  # functional.py:77-85
    return np.flip(img, axis)


def crop_3d(img, i, j, k, d, h, w):  # functional.py:61-74
    return img[i:i + d, j:j + h, k:k + w]


def resize_3d(img, size, order=1):  # functional.py:25-58 (3-D inputs)
    d, h, w = img.shape
    if isinstance(size, int):
        if min(d, h, w) == size:
            return img
        short = min(d, h, w)
        od, oh, ow = int(size * d / short), int(size * h / short), int(size * w / short)
    else:
        od, oh, ow = size[0], size[1], size[2]
    return po.zoom_to_shape(np.asarray(img), (od, oh, ow), order)


def resized_crop_3d(img, i, j, k, d, h, w, size, interpolation):  # functional.py:103-110
    return resize_3d(crop_3d(img, i, j, k, d, h, w), size, order=interpolation)


class Compose:  # transform.py:27-72
    def __init__(self, transforms):
        if not isinstance(transforms, list):
            raise TypeError("The transforms must be a list!")
        self.transforms = transforms

    def __call__(self, im, label=None):
        for op in self.transforms:
            res = op(im, label)
            im = res[0]
            if len(res) == 2:
                label = res[1]
        im = np.expand_dims(im, axis=0)
        if im.max() > 0:
            im = im / im.max()
        return im, label


class Resize3D:  # transform.py:74-109
    def __init__(self, size, order=1):
        self.size = size if isinstance(size, int) else tuple(size)
        self.order = order

    def __call__(self, img, label=None):
        img = resize_3d(img, self.size, self.order)
        if label is not None:
            label = resize_3d(label, self.size, 0)
        return img, label


class RandomRotation3D:  # transform.py:112-167 (label rotated with order 1 too, :163-165)
    def __init__(self, degrees, rotate_planes=((0, 1), (0, 2), (1, 2))):
        self.degrees = (-degrees, degrees) if isinstance(degrees, numbers.Number) else tuple(degrees)
        self.rotate_planes = [list(p) for p in rotate_planes]

    def __call__(self, img, label=None):
        angle = random.uniform(self.degrees[0], self.degrees[1])                       # :148
        plane = self.rotate_planes[random.randint(0, len(self.rotate_planes) - 1)]     # :149-150
        img = rotate_3d(img, plane, angle)
        if label is not None:
            label = rotate_3d(label, plane, angle)
        return img, label


class RandomFlip3D:  # transform.py:169-203
    def __init__(self, prob=0.5, flip_axis=(0, 1, 2)):
        self.prob, self.flip_axis = prob, flip_axis

    def __call__(self, img, label=None):
        if isinstance(self.flip_axis, (tuple, list)):
            axis = self.flip_axis[random.randint(0, len(self.flip_axis) - 1)]           # :194-195
        else:
            axis = self.flip_axis
        if random.random() < self.prob:                                               # :199
            img = flip_3d(img, axis)
            if label is not None:
                label = flip_3d(label, axis)
        return img, label


class RandomResizedCrop3D:  # transform.py:206-339 (pre_crop=False path, as the shipped configs use it)
    def __init__(self, size, scale=(0.8, 1.2), ratio=(3. / 4., 4. / 3.), interpolation=1):
        self.size = (size, size, size) if isinstance(size, int) else tuple(size)
        self.scale, self.ratio, self.interpolation = scale, ratio, interpolation

    def box(self, shape):  # get_params, transform.py:240-279
        for _ in range(10):
            target = random.uniform(*self.scale) * (shape[0] * shape[1] * shape[2])
            aspect = random.uniform(*self.ratio)
            d = int(round((target * aspect) ** (1 / 3)))
            h = int(round((target / aspect) ** (1 / 3)))
            w = shape[2]
            if random.random() < 0.5:
                d, h, w = random.sample([d, h, w], k=3)
            if w <= shape[2] and h <= shape[1] and d <= shape[0]:
                i = random.randint(0, shape[0] - d)
                j = random.randint(0, shape[1] - h)
                k = random.randint(0, shape[2] - w)
                return i, j, k, d, h, w
        side = min(shape)
        return (shape[0] - side) // 2, (shape[1] - side) // 2, (shape[2] - side) // 2, side, side, side

    def __call__(self, img, label=None):
        i, j, k, d, h, w = self.box(img.shape)
        img = resized_crop_3d(img, i, j, k, d, h, w, self.size, self.interpolation)
        if label is not None:
            label = resized_crop_3d(label, i, j, k, d, h, w, self.size, 0)
        return img, label
