"""GPU parity of the on-device augmentations (medicalseg_b200/transforms.py -> csrc/augment.cu, preprocess.cu through
the C ABI) against outputs of the reference's OWN transform code (tests/golden/transforms_ref.npz) and, at other
sizes, against the oracle restatement on the same seeded draws.

Tolerances: rotations and flips are bit-exact for labels and within one f32 rounding (1e-4 on the 0-255 scale) for
images - the kernel repeats SciPy's f64 coordinate arithmetic without FMA contraction; crop+zoom images within 1e-3
on the 0-255 scale (f32 interpolation in the resample kernel vs SciPy's f64), zoomed labels exact except at exact .5
coordinate ties (<= 0.2 % of voxels)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "transforms_ref.npz"))
PLANES = ([0, 1], [0, 2], [1, 2])


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("ai", range(7))
def test_rotation_matches_reference_outputs(ai):
    from medicalseg_b200 import transforms as T
    ang = float(G["rot_angles"][ai])
    img, lab = _dev(G["rot_img"]), _dev(G["rot_lab"])
    for pi, plane in enumerate(PLANES):
        out = T.rotate_3d(img, plane, ang).cpu().numpy()
        np.testing.assert_allclose(out, G["rot_img_%d_%d" % (ai, pi)], rtol=0, atol=1e-4)
        np.testing.assert_array_equal(T.rotate_3d(lab, plane, ang).cpu().numpy(), G["rot_lab_%d_%d" % (ai, pi)])
    if ai == 0:
        np.testing.assert_array_equal(T.rotate_3d(img, [1, 2], 28.4, order=0).cpu().numpy(), G["rot_img_order0"])
        np.testing.assert_allclose(T.rotate_3d(img, [0, 2], -41.0, cval=7).cpu().numpy(), G["rot_img_cval"], atol=1e-4)


def test_flip_crop_resize_match_reference_outputs():
    from medicalseg_b200 import transforms as T
    img, lab = _dev(G["rot_img"]), _dev(G["rot_lab"])
    for ax in range(3):
        np.testing.assert_array_equal(T.flip_3d(img, ax).cpu().numpy(), G["flip_%d" % ax])
    out = T.resized_crop_3d(img, 2, 3, 1, 9, 12, 8, (10, 10, 10), 1).cpu().numpy()
    np.testing.assert_allclose(out, G["rcrop_img"], rtol=0, atol=1e-3)
    outl = T.resized_crop_3d(lab, 2, 3, 1, 9, 12, 8, (10, 10, 10), 0).cpu().numpy()
    assert (outl != G["rcrop_lab"]).mean() <= 2e-3
    out = T.resize_3d(img, 8, 1).cpu().numpy()
    assert out.shape == G["resize_int"].shape
    np.testing.assert_allclose(out, G["resize_int"], rtol=0, atol=1e-3)
    a, b = T.Resize3D(size=[10, 12, 11])(G["pipe_img"], G["pipe_lab"])
    np.testing.assert_allclose(a.cpu().numpy(), G["cls_resize_img"], rtol=0, atol=1e-3)
    assert (b.cpu().numpy() != G["cls_resize_lab"]).mean() <= 2e-3


@pytest.mark.parametrize("seed", range(6))
def test_random_transform_classes_replay_the_reference(seed):
    """same `random` seed -> same angle / plane / flip / crop box as the reference classes -> same volumes"""
    from medicalseg_b200 import transforms as T
    img, lab = G["pipe_img"], G["pipe_lab"]
    random.seed(100 + seed); np.random.seed(100 + seed)
    a, b = T.RandomRotation3D(degrees=90)(img, lab)
    assert a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.int32
    np.testing.assert_allclose(a.cpu().numpy(), G["cls_rot_img_%d" % seed], rtol=0, atol=1e-4)
    np.testing.assert_array_equal(b.cpu().numpy(), G["cls_rot_lab_%d" % seed])
    random.seed(200 + seed); np.random.seed(200 + seed)
    a, b = T.RandomFlip3D()(img, lab)
    np.testing.assert_array_equal(a.cpu().numpy(), G["cls_flip_img_%d" % seed])
    np.testing.assert_array_equal(b.cpu().numpy(), G["cls_flip_lab_%d" % seed])
    random.seed(300 + seed); np.random.seed(300 + seed)
    a, b = T.RandomResizedCrop3D(size=16, scale=[0.8, 1.2])(img, lab)
    np.testing.assert_allclose(a.cpu().numpy(), G["cls_crop_img_%d" % seed], rtol=0, atol=1e-3)
    assert (b.cpu().numpy() != G["cls_crop_lab_%d" % seed]).mean() <= 2e-3
    # the lung_coronavirus.yml training pipeline (configs/lung_coronavirus/lung_coronavirus.yml:10-16)
    random.seed(400 + seed); np.random.seed(400 + seed)
    pipe = T.Compose([T.RandomResizedCrop3D(size=16, scale=[0.8, 1.2]), T.RandomRotation3D(degrees=90),
                      T.RandomFlip3D()])
    a, b = pipe(img, lab)
    assert tuple(a.shape) == (1, 16, 16, 16) and float(a.max()) == 1.0
    np.testing.assert_allclose(a.cpu().numpy(), G["pipe_out_img_%d" % seed], rtol=0, atol=1e-5)  # 0-1 scale
    # a zoom tie flips one source voxel; the order-1 label rotation can spread it over its 4 neighbours
    assert (b.cpu().numpy() != G["pipe_out_lab_%d" % seed]).mean() <= 5e-3


def test_full_size_properties_and_oracle_at_another_size():
    """128^3: quarter turns are exact permutations (degree-exact cos / sin), flips are involutions, Compose scales
    to max 1; 40x56x48 with a random angle agrees with the oracle."""
    from medicalseg_b200 import transforms as T
    from oracle import transforms_oracle as to
    g = torch.Generator().manual_seed(0)
    vol = torch.rand(128, 128, 128, generator=g).cuda() * 300
    lab = torch.randint(0, 5, (128, 128, 128), generator=g, dtype=torch.int32).cuda()
    for plane in PLANES:
        r = T.rotate_3d(vol, plane, 90.0)
        assert torch.equal(r, torch.rot90(vol, -1, dims=plane)) or torch.equal(r, torch.rot90(vol, 1, dims=plane))
        back = T.rotate_3d(r, plane, -90.0)
        assert torch.equal(back, vol)
        assert torch.equal(T.rotate_3d(T.rotate_3d(lab, plane, 180.0), plane, 180.0), lab)
    for ax in range(3):
        assert torch.equal(T.flip_3d(T.flip_3d(vol, ax), ax), vol)
        assert torch.equal(T.flip_3d(vol, ax), torch.flip(vol, [ax]))
    im, _ = T.Compose([])(vol, lab)
    assert tuple(im.shape) == (1, 128, 128, 128) and float(im.max()) == 1.0
    assert torch.equal(im[0], vol / vol.max())
    zero, _ = T.Compose([])(torch.zeros(4, 4, 4).cuda())
    assert float(zero.abs().max()) == 0.0  # max == 0: left unscaled (transform.py:68)
    rng = np.random.default_rng(5)
    img2 = (rng.random((40, 56, 48)) * 255).astype(np.float32)
    lab2 = rng.integers(0, 4, size=(40, 56, 48)).astype(np.int32)
    for plane, ang in zip(PLANES, (33.21, -77.7, 5.5)):
        np.testing.assert_allclose(T.rotate_3d(_dev(img2), plane, ang).cpu().numpy(), to.rotate_3d(img2, plane, ang),
                                   rtol=0, atol=1e-4)
        np.testing.assert_array_equal(T.rotate_3d(_dev(lab2), plane, ang).cpu().numpy(), to.rotate_3d(lab2, plane, ang))


def test_dataset_applies_the_configured_transforms_on_the_device(tmp_path):
    """datasets/dataset.py:113-118: the YAML `transforms` list is composed and applied per item"""
    from medicalseg_b200.cvlibs import Config
    root = tmp_path / "ds"
    (root / "images").mkdir(parents=True)
    (root / "labels").mkdir()
    rng = np.random.default_rng(0)
    with open(root / "train_list.txt", "w") as f:
        for i in range(2):
            np.save(root / "images" / ("v%d.npy" % i), (rng.random((24, 28, 20)) * 255).astype(np.float32))
            np.save(root / "labels" / ("v%d.npy" % i), rng.integers(0, 3, size=(24, 28, 20)).astype(np.uint8))
            f.write("images/v%d.npy labels/v%d.npy\n" % (i, i))
    cfg = tmp_path / "cfg.yml"
    cfg.write_text("""
batch_size: 2
iters: 2
train_dataset:
  type: LungCoronavirus
  dataset_root: %s
  result_dir: %s
  transforms:
    - type: RandomResizedCrop3D
      size: 16
      scale: [0.8, 1.2]
    - type: RandomRotation3D
      degrees: 90
    - type: RandomFlip3D
  mode: train
  num_classes: 3
model:
  type: VNet
  num_classes: 3
""" % (root, root))
    ds = Config(str(cfg)).train_dataset
    assert len(ds) == 20  # train lists are repeated 10x (dataset.py:110-111)
    random.seed(0)
    im, lab, path = ds[1]
    assert im.is_cuda and tuple(im.shape) == (1, 16, 16, 16) and im.dtype == torch.float32 and float(im.max()) == 1.0
    assert lab.is_cuda and tuple(lab.shape) == (16, 16, 16) and lab.dtype == torch.int32
    assert int(lab.min()) >= 0 and int(lab.max()) <= 2 and path.endswith("v1.npy")
