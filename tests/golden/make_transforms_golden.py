"""Generates tests/golden/transforms_ref.npz from the reference's OWN augmentation code.  Run in the BUILD container
only (needs /root/reference):

    python tests/golden/make_transforms_golden.py

medicalseg/transforms/transform.py and functional.py are imported by file path with two stubs: `SimpleITK` (only used
by the connected-component post-processing, not by the augmentations) and `medicalseg.cvlibs.manager` (the component
registry decorator).  Every case seeds Python's `random` (and numpy's) so the oracle and the device path can replay
the same parameter draws.
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_transforms():
    sys.modules["SimpleITK"] = types.ModuleType("SimpleITK")

    class _Mgr:
        def add_component(self, c):
            return c
    manager = types.ModuleType("medicalseg.cvlibs.manager")
    manager.TRANSFORMS = _Mgr()
    for name in ("medicalseg", "medicalseg.cvlibs", "medicalseg.transforms"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["medicalseg.cvlibs.manager"] = manager
    sys.modules["medicalseg.cvlibs"].manager = manager
    mods = {}
    for fn in ("functional", "transform"):
        spec = importlib.util.spec_from_file_location("medicalseg.transforms." + fn,
                                                      os.path.join(REF, "medicalseg/transforms", fn + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["medicalseg.transforms." + fn] = mod
        setattr(sys.modules["medicalseg.transforms"], fn, mod)
        spec.loader.exec_module(mod)
        mods[fn] = mod
    return mods["functional"], mods["transform"]


def volume(shape, seed, classes=4):
    rng = np.random.default_rng(seed)
    img = (rng.random(shape) * 255).astype(np.float32)
    lab = rng.integers(0, classes, size=shape).astype(np.int32)
    return img, lab


def main():
    F, T = load_reference_transforms()
    out = {}
    # 1. plane rotations (functional.rotate_3d): image (order 1) and label (order 1, the reference's quirk)
    img, lab = volume((12, 18, 10), 1)
    out["rot_img"], out["rot_lab"] = img, lab
    angles = [17.3, -63.9, 90.0, 180.0, -90.0, 0.0, 3.0001]
    out["rot_angles"] = np.asarray(angles)
    for ai, ang in enumerate(angles):
        for pi, plane in enumerate(([0, 1], [0, 2], [1, 2])):
            out["rot_img_%d_%d" % (ai, pi)] = F.rotate_3d(img, plane, ang)
            out["rot_lab_%d_%d" % (ai, pi)] = F.rotate_3d(lab, plane, ang)
    out["rot_img_order0"] = F.rotate_3d(img, [1, 2], 28.4, order=0)
    out["rot_img_cval"] = F.rotate_3d(img, [0, 2], -41.0, cval=7)
    # 2. flips, crops + zoom
    for ax in range(3):
        out["flip_%d" % ax] = np.ascontiguousarray(F.flip_3d(img, ax))
    out["rcrop_img"] = F.resized_crop_3d(img, 2, 3, 1, 9, 12, 8, (10, 10, 10), 1)
    out["rcrop_lab"] = F.resized_crop_3d(lab, 2, 3, 1, 9, 12, 8, (10, 10, 10), 0)
    out["resize_int"] = F.resize_3d(img, 8, 1)
    # 3. the classes with seeded draws, one at a time and as the lung_coronavirus.yml pipeline
    img, lab = volume((20, 24, 22), 2, classes=3)
    out["pipe_img"], out["pipe_lab"] = img, lab
    for seed in range(6):
        random.seed(100 + seed); np.random.seed(100 + seed)
        a, b = T.RandomRotation3D(degrees=90)(img, lab)
        out["cls_rot_img_%d" % seed], out["cls_rot_lab_%d" % seed] = a, b
        random.seed(200 + seed); np.random.seed(200 + seed)
        a, b = T.RandomFlip3D()(img, lab)
        out["cls_flip_img_%d" % seed], out["cls_flip_lab_%d" % seed] = np.ascontiguousarray(a), np.ascontiguousarray(b)
        random.seed(300 + seed); np.random.seed(300 + seed)
        a, b = T.RandomResizedCrop3D(size=16, scale=[0.8, 1.2])(img, lab)
        out["cls_crop_img_%d" % seed], out["cls_crop_lab_%d" % seed] = a, b
        random.seed(400 + seed); np.random.seed(400 + seed)
        pipe = T.Compose([T.RandomResizedCrop3D(size=16, scale=[0.8, 1.2]), T.RandomRotation3D(degrees=90),
                          T.RandomFlip3D()])
        a, b = pipe(img, lab)
        out["pipe_out_img_%d" % seed], out["pipe_out_lab_%d" % seed] = np.ascontiguousarray(a), np.ascontiguousarray(b)
    a, b = T.Resize3D(size=[10, 12, 11])(img, lab)
    out["cls_resize_img"], out["cls_resize_lab"] = a, b
    np.savez_compressed(os.path.join(HERE, "transforms_ref.npz"), **out)
    print("transforms_ref.npz: %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(HERE, "transforms_ref.npz")) / 1e3))


if __name__ == "__main__":
    main()
