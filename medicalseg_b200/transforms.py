"""Training augmentations with the reference's class names, arguments and random-number draws, executed on the GPU
(reference: medicalseg/transforms/transform.py:27-339, functional.py:25-110).

The reference augments on the host with NumPy / SciPy inside DataLoader workers: `scipy.ndimage.rotate` of one 128^3
volume costs ~1 s of a core, ~80 train steps of this engine, so the data path would throttle training by two orders
of magnitude.  Here a sample is moved to the device once and every transform is one HBM-bound kernel
(csrc/augment.cu, csrc/preprocess.cu).  The random PARAMETERS are drawn on the host from Python's `random` / `numpy.random`
in the reference's order (same seed -> same crop boxes, angles, planes and flips); the resampling itself follows
SciPy's coordinate arithmetic (see tests/test_gpu_transforms.py for the parity statement).

Inputs: numpy arrays or torch tensors [D,H,W]; outputs stay on the device (float32 image, int32 label)."""
from __future__ import annotations

import collections.abc
import numbers
import random

import numpy as np
import torch

from . import ops
from . import preprocess as P


def _dev_pair(im, label):
    dev = P._device()
    if isinstance(im, str):
        im = np.load(im)
    if isinstance(label, str):
        label = np.load(label)
    if im is None:
        raise ValueError("Can't read The image file {}!".format(im))
    if not torch.is_tensor(im):
        im = torch.from_numpy(np.ascontiguousarray(im))
    im = im.to(dev).to(torch.float32).contiguous()
    if label is not None:
        if not torch.is_tensor(label):
            label = torch.from_numpy(np.ascontiguousarray(label))
        label = label.to(dev).to(torch.int32).contiguous()
    return im, label


def rotation_coefficients(plane_shape, angle):
    """(M, offset) of scipy.ndimage.rotate(reshape=False) for a plane of `plane_shape` = (n_a, n_b): input coordinate
    = M @ output coordinate + offset, with the degree-exact cosine / sine SciPy uses (cosdg / sindg)."""
    try:
        from scipy import special
        c, s = float(special.cosdg(angle)), float(special.sindg(angle))
    except ImportError:  # same values except in the last ulp away from multiples of 90 degrees
        import math
        q = (angle / 90.0) % 4
        exact = {0.0: (1.0, 0.0), 1.0: (0.0, 1.0), 2.0: (-1.0, 0.0), 3.0: (0.0, -1.0)}
        c, s = exact.get(q, (math.cos(math.radians(angle)), math.sin(math.radians(angle))))
    m = np.array([[c, s], [-s, c]])
    center = (np.asarray(plane_shape, dtype=np.float64) - 1) / 2
    off = center - m @ center
    return (m[0, 0], m[0, 1], m[1, 0], m[1, 1]), (float(off[0]), float(off[1]))


def rotate_3d(img, r_plane, angle, order=1, cval=0):
    """functional.py:91-100 on the device (image f32 -> f32; integer label -> i32 with SciPy's rounding)"""
    a, b = sorted(int(x) % 3 for x in r_plane)
    m, off = rotation_coefficients((img.shape[a], img.shape[b]), angle)
    out = torch.empty_like(img)
    ops.rotate3d(img, out, a, b, m, off, order, cval)
    return out


def flip_3d(img, axis):
    out = torch.empty_like(img)
    ops.flip3d(img, out, int(axis) % 3)
    return out


def resize_3d(img, size, order=1):
    """functional.py:25-58: scipy.ndimage.zoom(mode='nearest') to `size` (int = shortest side) on the device"""
    d, h, w = img.shape
    if isinstance(size, int):
        if min(d, h, w) == size:
            return img
        od, oh, ow = int(size * d / min(d, h, w)), int(size * h / min(d, h, w)), int(size * w / min(d, h, w))
    elif isinstance(size, collections.abc.Iterable) and len(size) == 3:
        od, oh, ow = (int(s) for s in size)
    else:
        raise TypeError("Got inappropriate size arg: {}".format(size))
    if order not in (0, 1):
        raise NotImplementedError("device resize supports interpolation order 0 and 1 (the shipped configs use 1)")
    return P.resample(img.contiguous(), new_shape=[od, oh, ow], order=order)[0]


def resized_crop_3d(img, i, j, k, d, h, w, size, interpolation):
    return resize_3d(img[i:i + d, j:j + h, k:k + w].contiguous(), size, order=interpolation)


class Compose:
    """transform.py:27-72: applies the transforms, adds the channel axis and divides by the volume maximum."""

    def __init__(self, transforms):
        if not isinstance(transforms, list):
            raise TypeError("The transforms must be a list!")
        self.transforms = transforms

    def __call__(self, im, label=None):
        im, label = _dev_pair(im, label)
        for op in self.transforms:
            outputs = op(im, label)
            im = outputs[0]
            if len(outputs) == 2:
                label = outputs[1]
        im = im.contiguous()
        mm = torch.empty(2, dtype=torch.float32, device=im.device)
        out = torch.empty_like(im)
        ops.minmax(im, mm)
        ops.scale_by_max(im, out, mm)  # im / im.max() if im.max() > 0 - the maximum never visits the host
        return out.unsqueeze(0), label


class Resize3D:
    def __init__(self, size, order=1):
        if isinstance(size, int):
            self.size = size
        elif isinstance(size, collections.abc.Iterable) and len(size) == 3:
            self.size = tuple(size)
        else:
            raise ValueError("Unknown inputs for size: {}".format(size))
        self.order = order

    def __call__(self, img, label=None):
        img, label = _dev_pair(img, label)
        img = resize_3d(img, self.size, self.order)
        if label is not None:
            label = resize_3d(label, self.size, 0)
        return img, label


class RandomRotation3D:
    """transform.py:112-167.  The label is rotated with the SAME interpolation order as the image (order 1 + SciPy's
    integer rounding) - a quirk of the reference that is kept."""

    def __init__(self, degrees, rotate_planes=[[0, 1], [0, 2], [1, 2]]):
        if isinstance(degrees, numbers.Number):
            if degrees < 0:
                raise ValueError("If degrees is a single number, it must be positive.")
            self.degrees = (-degrees, degrees)
        else:
            if len(degrees) != 2:
                raise ValueError("If degrees is a sequence, it must be of len 2.")
            self.degrees = degrees
        self.rotate_planes = rotate_planes

    def get_params(self, degrees):
        angle = random.uniform(degrees[0], degrees[1])
        r_plane = self.rotate_planes[random.randint(0, len(self.rotate_planes) - 1)]
        return angle, r_plane

    def __call__(self, img, label=None):
        img, label = _dev_pair(img, label)
        angle, r_plane = self.get_params(self.degrees)
        img = rotate_3d(img, r_plane, angle)
        if label is not None:
            label = rotate_3d(label, r_plane, angle)
        return img, label


class RandomFlip3D:
    """transform.py:169-203"""

    def __init__(self, prob=0.5, flip_axis=[0, 1, 2]):
        self.prob, self.flip_axis = prob, flip_axis

    def __call__(self, img, label=None):
        img, label = _dev_pair(img, label)
        if isinstance(self.flip_axis, (tuple, list)):
            flip_axis = self.flip_axis[random.randint(0, len(self.flip_axis) - 1)]
        else:
            flip_axis = self.flip_axis
        if random.random() < self.prob:
            img = flip_3d(img, flip_axis)
            if label is not None:
                label = flip_3d(label, flip_axis)
        return img, label


class RandomResizedCrop3D:
    """transform.py:206-339: crop a random box (volume ratio `scale`, aspect jitter `ratio`), zoom it to `size`;
    label with order 0.  `pre_crop` / `nonzero_mask` as in the reference (the mask bounds are read back once)."""

    def __init__(self, size, scale=(0.8, 1.2), ratio=(3. / 4., 4. / 3.), interpolation=1, pre_crop=False,
                 nonzero_mask=False):
        if isinstance(size, (tuple, list)):
            assert len(size) == 3, \
                "Size must contain THREE number when it is a tuple or list, got {}.".format(len(size))
            self.size = size
        elif isinstance(size, int):
            self.size = (size, size, size)
        else:
            raise TypeError("Size must be a list or tuple, got {}.".format(type(size)))
        self.interpolation, self.scale, self.ratio = interpolation, scale, ratio
        self.pre_crop, self.nonzero_mask = pre_crop, nonzero_mask

    def get_params(self, img, scale, ratio):
        shape = tuple(img.shape)
        for _ in range(10):
            volume = shape[0] * shape[1] * shape[2]
            target_volume = random.uniform(*scale) * volume
            aspect_ratio = random.uniform(*ratio)
            d = int(round((target_volume * aspect_ratio) ** (1 / 3)))
            h = int(round((target_volume / aspect_ratio) ** (1 / 3)))
            w = shape[2]
            if random.random() < 0.5:
                d, h, w = random.sample([d, h, w], k=3)
            if w <= shape[2] and h <= shape[1] and d <= shape[0]:
                i = random.randint(0, shape[0] - d)
                j = random.randint(0, shape[1] - h)
                k = random.randint(0, shape[2] - w)
                return i, j, k, d, h, w
        w = min(shape)  # fallback: central cube
        return (shape[0] - w) // 2, (shape[1] - w) // 2, (shape[2] - w) // 2, w, w, w

    def pre_crop_util(self, img, label=None):
        if not self.pre_crop:
            return img, label
        crop_size = (np.random.uniform(low=self.scale[0], high=self.scale[1], size=3) * self.size).round().astype("int")
        if self.nonzero_mask:
            nz = torch.nonzero(label != 0)
            lo, hi = nz.min(0).values.tolist(), (nz.max(0).values + 1).tolist()
            masked_shape = np.array([hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]])
            crop = np.minimum(masked_shape, crop_size)
            start = [int(np.random.randint(masked_shape[a] - crop[a] + 1)) + lo[a] for a in range(3)]
        else:
            crop = np.minimum(np.array(img.shape[:3]), crop_size)
            start = [int(np.random.randint(img.shape[a] - crop[a] + 1)) for a in range(3)]
        sl = tuple(slice(start[a], start[a] + int(crop[a])) for a in range(3))
        return img[sl], (label[sl] if label is not None else None)

    def __call__(self, img, label=None):
        img, label = _dev_pair(img, label)
        img, label = self.pre_crop_util(img, label)
        i, j, k, d, h, w = self.get_params(img, self.scale, self.ratio)
        img = resized_crop_3d(img, i, j, k, d, h, w, self.size, self.interpolation)
        if label is not None:
            label = resized_crop_3d(label, i, j, k, d, h, w, self.size, 0)
        return img, label
