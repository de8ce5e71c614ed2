"""Config / registry semantics of the reference (cvlibs/config.py, manager.py) on the torch host layer (CPU), plus a
GPU smoke test of the train.py / val.py entry points on the synthetic configuration."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_base_inheritance_and_cli_overrides():
    from medicalseg_b200.cvlibs import Config
    cfg = Config(os.path.join(ROOT, "configs/lung_coronavirus/vnet_lung_coronavirus_128_128_128_15k.yml"),
                 learning_rate=0.01, iters=100, batch_size=4)
    assert cfg.dic["model"]["type"] == "VNet" and cfg.dic["model"]["num_classes"] == 3
    assert cfg.dic["optimizer"] == {"type": "sgd", "momentum": 0.9, "weight_decay": 1.0e-4}  # from the _base_ file
    assert cfg.dic["lr_scheduler"]["learning_rate"] == 0.01 and cfg.iters == 100 and cfg.batch_size == 4
    sched = cfg.lr_scheduler
    assert sched.decay_steps == 15000 and sched.power == 0.9 and sched.get_lr() == 0.01
    losses = cfg.loss
    assert type(losses["types"][0]).__name__ == "MixedLoss" and losses["coef"] == [1]
    inner = losses["types"][0].losses
    assert [type(l).__name__ for l in inner] == ["CrossEntropyLoss", "DiceLoss"] and inner[0].ignore_index == 255


def test_mri_config_resolves_fixed_base():
    from medicalseg_b200.cvlibs import Config
    cfg = Config(os.path.join(ROOT, "configs/mri_spine_seg/vnet_mri_spine_seg_512_512_12_15k.yml"))
    assert cfg.dic["model"]["kernel_size"][0] == [2, 2, 4] and cfg.dic["model"]["stride_size"][1] == [2, 2, 1]
    assert cfg.dic["lr_scheduler"]["learning_rate"] == 0.1 and cfg.dic["model"]["num_classes"] == 20


def test_config_errors_match_reference(tmp_path):
    from medicalseg_b200.cvlibs import Config, ComponentManager
    with pytest.raises(ValueError):
        Config("")
    with pytest.raises(FileNotFoundError):
        Config(str(tmp_path / "missing.yml"))
    bad = tmp_path / "c.txt"
    bad.write_text("a: 1")
    with pytest.raises(RuntimeError):
        Config(str(bad))
    y = tmp_path / "c.yml"
    y.write_text("batch_size: 2\nloss:\n  types:\n    - type: DiceLoss\n    - type: DiceLoss\n  coef: [1, 1, 1]\n")
    with pytest.raises(ValueError):
        Config(str(y)).loss
    with pytest.raises(RuntimeError):
        Config(str(y)).iters
    y.write_text("iters: 5\n_inherited_: true\nloss:\n  types:\n    - type: DiceLoss\n  coef: [0.5, 0.5]\n")
    l = Config(str(y)).loss  # a single type is broadcast over coef (config.py:259-262)
    assert len(l["types"]) == 2
    m = ComponentManager("x")

    class A:
        pass
    m.add_component(A)
    with pytest.raises(KeyError):
        m.add_component(A)
    with pytest.raises(KeyError):
        m["B"]
    with pytest.raises(TypeError):
        m.add_component(3)


def test_synthetic_dataset_contract():
    from medicalseg_b200.datasets import SyntheticVolumes
    ds = SyntheticVolumes(num_classes=3, shape=(16, 24, 32), length=4)
    im, lab, path = ds[1]
    assert tuple(im.shape) == (1, 16, 24, 32) and im.dtype == torch.float32 and float(im.max()) == 1.0
    assert tuple(lab.shape) == (16, 24, 32) and lab.dtype == torch.int32 and set(lab.unique().tolist()) <= {0, 1, 2}
    im2, _, _ = ds[1]
    assert torch.equal(im, im2)


def test_npy_dataset_reads_reference_file_lists(tmp_path):
    import numpy as np
    from medicalseg_b200.datasets import NpyVolumeDataset
    np.save(tmp_path / "a.npy", np.random.rand(4, 5, 6).astype(np.float32) * 200)
    np.save(tmp_path / "b.npy", np.random.randint(0, 3, (4, 5, 6)))
    (tmp_path / "train_list.txt").write_text("a.npy b.npy\n")
    ds = NpyVolumeDataset(str(tmp_path), num_classes=3, mode="train")
    assert len(ds) == 10  # the train list is replicated x10 (datasets/dataset.py:110-111)
    im, lab, p = ds[3]
    assert tuple(im.shape) == (1, 4, 5, 6) and abs(float(im.max()) - 1.0) < 1e-6 and lab.dtype == torch.int32
    with pytest.raises(ValueError):
        NpyVolumeDataset(str(tmp_path), num_classes=None)


@pytest.mark.gpu
def test_train_and_val_entry_points_run_on_synthetic_config(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    cfg = os.path.join(ROOT, "configs/synthetic/vnet_synthetic_64.yml")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", cfg, "--iters", "6", "--log_iters", "2",
                        "--save_interval", "6", "--do_eval", "--save_dir", str(tmp_path / "out"), "--seed", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[TRAIN] epoch:" in r.stdout and "ips:" in r.stdout and "[EVAL] #Images: 2, Dice:" in r.stdout
    assert os.path.exists(tmp_path / "out" / "iter_6" / "model.pdparams")
    assert os.path.exists(tmp_path / "out" / "iter_6" / "model.pdopt")
    assert os.path.exists(tmp_path / "out" / "best_model" / "model.pdparams")
    losses = [float(l.split("loss: ")[1].split(",")[0]) for l in r.stdout.splitlines() if "[TRAIN]" in l]
    assert losses[-1] < losses[0]  # it learns
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "val.py"), "--config", cfg, "--model_path",
                         str(tmp_path / "out" / "iter_6"), "--save_dir", str(tmp_path / "val")],
                        capture_output=True, text=True, env=env, timeout=600)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "[EVAL] #Images: 2, Dice:" in r2.stdout
    r3 = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", cfg, "--iters", "8", "--log_iters", "1",
                         "--save_interval", "100", "--resume_model", str(tmp_path / "out" / "iter_6"),
                         "--save_dir", str(tmp_path / "out2")], capture_output=True, text=True, env=env, timeout=600)
    assert r3.returncode == 0, r3.stderr[-2000:]
    assert "iter: 7/8" in r3.stdout and "iter: 6/8" not in r3.stdout  # resumed at iteration 6


@pytest.mark.gpu
def test_train_entry_point_with_to_static_training_cuda_graph(tmp_path):
    """--to_static_training (reference flag, train.py:88) captures the train step into one CUDA graph; the run must log,
    evaluate (eager forward with lazily re-packed weights) and checkpoint like the eager one, and learn."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    cfg = os.path.join(ROOT, "configs/synthetic/vnet_synthetic_64.yml")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", cfg, "--iters", "10", "--log_iters", "2",
                        "--save_interval", "10", "--do_eval", "--save_dir", str(tmp_path / "out"), "--seed", "0",
                        "--to_static_training"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[TRAIN] epoch:" in r.stdout and "[EVAL] #Images: 2, Dice:" in r.stdout
    assert os.path.exists(tmp_path / "out" / "iter_10" / "model.pdparams")
    assert os.path.exists(tmp_path / "out" / "best_model" / "model.pdparams")
    losses = [float(l.split("loss: ")[1].split(",")[0]) for l in r.stdout.splitlines() if "[TRAIN]" in l]
    assert len(losses) >= 3 and losses[-1] < losses[0]


@pytest.mark.gpu
def test_train_entry_point_deep_supervision_and_device_transforms(tmp_path):
    """VNetDeepSup through train.py (four losses, the heads' gradient bucket fires first - the reducer's planner checks
    the end-to-front order), then a .npy dataset whose YAML `transforms` run on the device (datasets/dataset.py:113)."""
    import numpy as np
    env = dict(os.environ, PYTHONPATH=ROOT)
    cfg = os.path.join(ROOT, "configs/synthetic/vnetdeepsup_synthetic_64.yml")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", cfg, "--iters", "6", "--log_iters", "2",
                        "--save_interval", "6", "--do_eval", "--save_dir", str(tmp_path / "out"), "--seed", "0",
                        "--to_static_training"],  # the four-output step is captured into one CUDA graph as well
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[EVAL] #Images: 2, Dice:" in r.stdout
    losses = [float(l.split("loss: ")[1].split(",")[0]) for l in r.stdout.splitlines() if "[TRAIN]" in l]
    assert len(losses) >= 2 and losses[-1] < losses[0]  # (the capture's 3 warm-up steps are not logged)
    root = tmp_path / "ds"
    root.mkdir()
    rng = np.random.default_rng(0)
    lines = []
    for i in range(2):
        blob = rng.random((40, 44, 36)).astype(np.float32)
        np.save(root / ("im%d.npy" % i), blob * 255)
        np.save(root / ("lb%d.npy" % i), (blob > 0.5).astype(np.uint8))
        lines.append("im%d.npy lb%d.npy" % (i, i))
    (root / "train_list.txt").write_text("\n".join(lines) + "\n")
    (root / "val_list.txt").write_text(lines[0] + "\n")
    cfg2 = tmp_path / "npy.yml"
    cfg2.write_text("""
_base_: '%s'
train_dataset:
  _inherited_: False
  type: LungCoronavirus
  dataset_root: %s
  result_dir: %s
  transforms:
    - type: RandomResizedCrop3D
      size: 32
      scale: [0.8, 1.2]
    - type: RandomRotation3D
      degrees: 90
    - type: RandomFlip3D
  mode: train
  num_classes: 2
val_dataset:
  _inherited_: False
  type: LungCoronavirus
  dataset_root: %s
  result_dir: %s
  transforms:
    - type: Resize3D
      size: [32, 32, 32]
  mode: val
  num_classes: 2
""" % (os.path.join(ROOT, "configs/synthetic/vnet_synthetic_64.yml"), root, root, root, root))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", str(cfg2), "--iters", "4",
                        "--log_iters", "2", "--save_interval", "4", "--do_eval", "--save_dir", str(tmp_path / "out2"),
                        "--seed", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[TRAIN] epoch:" in r.stdout and "[EVAL] #Images: 1, Dice:" in r.stdout


def test_shipped_configs_build_their_transform_pipelines():
    """the lung / MRI recipes list the reference's augmentations (lung_coronavirus.yml:10-16,
    mri_spine_seg_1e-1_big_rmresizecrop_class20.yml:10-13); the deep-supervision config inherits the VNet one and
    repeats its loss entry four times.  Host-side construction only (the datasets themselves are not shipped)."""
    from medicalseg_b200.cvlibs import Config
    from medicalseg_b200 import transforms as T
    lung = Config(os.path.join(ROOT, "configs/lung_coronavirus/vnet_lung_coronavirus_128_128_128_15k.yml"))
    tl = [lung._load_object(t) for t in lung.dic["train_dataset"]["transforms"]]
    assert [type(t) for t in tl] == [T.RandomResizedCrop3D, T.RandomRotation3D, T.RandomFlip3D]
    assert tl[0].size == (128, 128, 128) and tl[0].scale == [0.8, 1.2] and tl[1].degrees == (-90, 90)
    assert lung.dic["val_dataset"]["transforms"] == []
    mri = Config(os.path.join(ROOT, "configs/mri_spine_seg/vnetdeepsup_mri_spine_seg_512_512_12_15k.yml"))
    tm = [mri._load_object(t) for t in mri.dic["train_dataset"]["transforms"]]
    assert [type(t) for t in tm] == [T.RandomRotation3D, T.RandomFlip3D] and tm[0].degrees == (-30, 30)
    assert mri.dic["model"]["type"] == "VNetDeepSup" and mri.dic["model"]["num_classes"] == 20
    assert mri.dic["model"]["kernel_size"][0] == [2, 2, 4] and mri.dic["model"]["stride_size"][1] == [2, 2, 1]
    losses = mri.loss
    assert len(losses["types"]) == 4 and losses["coef"] == [0.25] * 4
    assert len({id(l) for l in losses["types"]}) == 4  # four separate objects (each caches its own CE class weights)
