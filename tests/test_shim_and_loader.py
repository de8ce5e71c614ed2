"""CPU tests of round-2 host logic: the `medicalseg` import shim (reference package names resolve to the B200
implementation), the distributed batch sampler (Paddle's padding contract), the restricted checkpoint loader and the
direct NCCL binding's symbol table."""
import os
import pickle

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve_to_the_b200_implementation():
    """medicalseg/models/__init__.py:15-17, cvlibs/__init__.py:15-16, core/__init__.py:15-17, train.py:21-23, val.py:20-22"""
    import medicalseg
    from medicalseg.models import VNet, VNetDeepSup, DiceLoss, CrossEntropyLoss, MixedLoss
    from medicalseg.cvlibs import manager, Config
    from medicalseg.core import train, evaluate, infer
    from medicalseg.utils import get_sys_env, logger, config_check, utils, loss_computation
    from medicalseg.transforms import Compose, RandomFlip3D, RandomResizedCrop3D, RandomRotation3D, Resize3D
    from medicalseg.datasets import MedicalDataset, LungCoronavirus, MRISpineSeg
    import medicalseg_b200.models as M
    import medicalseg_b200.core as C
    assert VNet is M.VNet and VNetDeepSup is M.VNetDeepSup and DiceLoss is M.DiceLoss
    assert train is C.train and evaluate is C.evaluate and infer.inference is C.inference
    # YAML `type: VNet` goes through the same registry object user code decorates (manager.py:93-149)
    assert manager.MODELS["VNet"] is VNet and manager.LOSSES["MixedLoss"] is MixedLoss
    assert manager.TRANSFORMS["Resize3D"] is Resize3D and manager.DATASETS["LungCoronavirus"] is LungCoronavirus

    @manager.MODELS.add_component
    class _UserModel:  # a reference-style user extension
        pass
    assert manager.MODELS["_UserModel"] is _UserModel
    cfg = Config(os.path.join(ROOT, "configs/lung_coronavirus/vnet_lung_coronavirus_128_128_128_15k.yml"))
    assert cfg.dic["model"]["type"] == "VNet"
    assert isinstance(get_sys_env(), dict) and callable(logger.info) and callable(utils.load_entire_model)
    with pytest.raises(NotImplementedError):
        medicalseg.models.losses.BCELoss


def test_config_check_matches_reference_rules(tmp_path):
    from medicalseg.utils import config_check

    class DS:
        def __init__(self, n):
            self.num_classes = n

    class Cfg:
        def __init__(self, model_nc, has_ds=True):
            self.dic = {"model": {"type": "VNet", "num_classes": model_nc}}
            self.train_dataset = self.val_dataset = object() if has_ds else None

    a, b = DS(3), DS(3)
    config_check(Cfg(3), a, b)
    assert a.num_classes == 3
    with pytest.raises(ValueError, match="not consistent"):
        config_check(Cfg(4), DS(3), None)
    with pytest.raises(ValueError, match="should be given"):
        config_check(Cfg(3, has_ds=False), None, None)


@pytest.mark.parametrize("n,world,bs", [(10, 4, 2), (3, 8, 2), (16, 2, 2), (7, 2, 4), (200, 8, 2), (5, 1, 2)])
def test_distributed_batch_sampler_pads_like_paddle(n, world, bs):
    """paddle.io.DistributedBatchSampler (core/train.py:87-89): every rank gets ceil(n/world) samples in the same
    number of equally shaped batches; the union covers the dataset; padding repeats leading indices."""
    from medicalseg_b200.datasets import DistributedBatchSampler
    eps = [DistributedBatchSampler(n, bs, r, world, seed=3).epoch() for r in range(world)]
    shapes = [[len(b) for b in e] for e in eps]
    assert all(s == shapes[0] for s in shapes), shapes
    per_rank = -(-n // world)
    assert all(sum(s) == per_rank for s in shapes)
    flat = [i for e in eps for b in e for i in b]
    assert set(flat) == set(range(n)) and len(flat) == per_rank * world
    assert len(DistributedBatchSampler(n, bs, 0, world)) == len(eps[0])
    # the stream never ends and reshuffles between epochs
    s = DistributedBatchSampler(n, bs, 0, world, seed=3)
    it = iter(s)
    got = [next(it) for _ in range(3 * len(s))]
    assert len(got) == 3 * len(s)
    with pytest.raises(ValueError):
        DistributedBatchSampler(0, 2)


def test_checkpoint_loader_reads_numpy_containers_and_rejects_code(tmp_path):
    from medicalseg_b200.utils import _load_any
    good = tmp_path / "model.pdparams"
    sd = {"w": np.arange(6, dtype=np.float32).reshape(2, 3), "StructuredToParameterName@@": {"w": "w"}}
    with open(good, "wb") as fh:
        pickle.dump(sd, fh, protocol=2)
    out = _load_any(str(good))
    np.testing.assert_array_equal(out["w"], sd["w"])
    assert _load_any(str(tmp_path))["w"].shape == (2, 3)  # a directory resolves to <dir>/model.pdparams
    t = tmp_path / "t.pdparams"
    torch.save({"w": torch.ones(3)}, t)
    assert torch.equal(_load_any(str(t))["w"], torch.ones(3))

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))
    bad = tmp_path / "evil.pdparams"
    with open(bad, "wb") as fh:
        pickle.dump({"w": Evil()}, fh, protocol=2)
    with pytest.raises(pickle.UnpicklingError):
        _load_any(str(bad))
    assert not (tmp_path / "pwned").exists()


def test_direct_nccl_binding_loads_the_library_torch_links():
    from medicalseg_b200 import nccl
    v = nccl.version()
    major, minor = torch.cuda.nccl.version()[:2] if hasattr(torch.cuda, "nccl") else (v // 10000, 0)
    assert v // 10000 == major and v >= 20906  # graph capture of collectives needs NCCL >= 2.9.6
    with pytest.raises(RuntimeError):
        nccl.Communicator(torch.device("cpu"))  # needs an initialised torch.distributed group for the rendezvous


def test_in_channels_must_divide_16_like_the_reference_tile():
    """vnet.py:74-79: x.tile(16 / in_channels) + conv output; anything else cannot add up to 16 channels"""
    from medicalseg_b200.models import VNet
    with pytest.raises(ValueError):
        VNet(in_channels=3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):  # valid arguments, but no CUDA device: no CPU fallback
            VNet(in_channels=2)


def test_inference_reverse_list_follows_resize_transforms():
    from medicalseg_b200.core import get_reverse_list

    class Resize3D:
        def __init__(self, size):
            self.size = size

    class Other:
        pass
    assert get_reverse_list((10, 20, 30), [Other(), Resize3D((4, 5, 6)), Resize3D((2, 2, 2))]) == [
        ("resize", (10, 20, 30)), ("resize", (4, 5, 6))]
    assert get_reverse_list((1, 2, 3), None) == []
