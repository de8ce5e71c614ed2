"""CPU tests: the C-ABI library loads and exports every symbol include/medseg_b200.h declares (no compute calls
without a GPU), argument validation works without touching the device, and the host-side logic (parameter store,
LR schedule, bucket planner, loss/optimizer argument errors) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "medseg_b200.h")


@pytest.fixture(scope="module")
def lib():
    from medicalseg_b200 import build, _lib
    build.build_library()  # nvcc cross-compiles sm_100a without a GPU; no-op when up to date
    return _lib


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 35
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), "symbol %s declared in the header but not exported" % n
        assert n in lib.SIGNATURES, "symbol %s has no ctypes signature" % n
    assert set(lib.SIGNATURES) == set(names)
    assert lib.call("msb_version") == 100


def test_signature_arity_matches_header(lib):
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_argument_validation_reports_errors_without_a_gpu(lib):
    with pytest.raises(lib.MsbError, match="msb_momentum_step"):
        lib.call("msb_momentum_step", None, None, None, 0, 0.1, 0.9, 0.0, 1.0, None)
    with pytest.raises(lib.MsbError, match="C <= 32"):
        lib.call("msb_dice_ce_fwd", None, None, None, 1, 64, 10, 255, None, None)
    with pytest.raises(lib.MsbError, match="order"):
        lib.call("msb_resample_f32", 1, lib.MsbDim3(2, 2, 2), 1, lib.MsbDim3(2, 2, 2), 3, 0, 0.0, 0.0, 0.0, None)
    t = lib.MsbTensor(None, 0, 12, 0)  # 12 channels: not a multiple of 8
    with pytest.raises(lib.MsbError):
        lib.call("msb_bn_stats", t, 1, 8, 1, None, None)
    assert lib.call("msb_conv_k5_packed_bytes", 32, 32) == 32 * 125 * 32 * 2
    assert lib.call("msb_conv_k5_out_pad", 24) == 32
    assert lib.call("msb_conv_k5_wgrad_workspace_bytes", 32, 2) == 125 * 32 * 2 * 4


def test_missing_library_fails_loudly(monkeypatch, lib):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libmedseg_b200.so")
    with pytest.raises(lib.MsbError, match="no CPU/PyTorch fallback"):
        lib.load()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_has_no_cpu_fallback():
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200 import preprocess as P
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VNet(num_classes=2)
    with pytest.raises(RuntimeError):
        L.DiceLoss()(torch.rand(1, 2, 4, 4, 4), torch.zeros(1, 4, 4, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.HUnorm(np.zeros((4, 4, 4), np.float32))


def test_polynomial_decay_matches_oracle():
    from medicalseg_b200.optimizer import PolynomialDecay
    from oracle import vnet_oracle as vo
    a, b = PolynomialDecay(0.001, 15000, end_lr=0, power=0.9), vo.PolynomialDecay(0.001, 15000)
    for _ in range(50):
        assert a.get_lr() == b.get_lr()
        a.step(); b.step()
    a.last_epoch = 20000
    assert a.get_lr() == 0.0


def test_param_store_layout_matches_reference_parameter_tree():
    """flat layout: forward order, 16-byte aligned slots, buffers separate, spans contiguous per block"""
    from medicalseg_b200.models import vnet as V
    from oracle import vnet_oracle as vo

    class Eng:  # minimal engine stand-in: only the store is needed to build the tree
        def register_packer(self, packer):
            pass
    eng = Eng()
    eng.store = V.ParamStore()
    eng.dtype = torch.float32
    blocks = [V.InputTransition(eng, "in_tr", 1), V.DownTransition(eng, "down_tr32", 16, 1, False, (2, 2, 2), (2, 2, 2)),
              V.UpTransition(eng, "up_tr32", 64, 32, 1, False, False, (2, 2, 2), (2, 2, 2)),
              V.OutputTransition(eng, "out_tr", 32, 3)]
    eng.store.finalize("cpu")
    om = vo.VNetOracle(num_classes=3)
    osd = om.state_dict()
    for name, slot in eng.store.slots.items():
        assert name in osd and tuple(osd[name].shape) == slot.shape, name
        assert slot.offset % 4 == 0
    prev_hi = 0
    for b in blocks:
        lo, hi = eng.store.span(b.all_param_names())
        assert lo == prev_hi  # blocks tile the flat buffer in forward order
        prev_hi = hi
    assert prev_hi == eng.store.flat.numel()
    # zero-padded physical slots for the 16-channel-padded head
    assert eng.store.phys("out_tr.bn1.weight").numel() == 16 and eng.store.view("out_tr.bn1.weight").numel() == 3


def test_bucket_planner():
    from medicalseg_b200.parallel import BucketPlanner
    bp = BucketPlanner(100, 30)
    assert bp.add(90, 100) is None
    assert bp.add(80, 90) is None
    assert bp.add(60, 80) == (60, 100)
    assert bp.add(50, 60) is None
    assert bp.add(0, 50) == (0, 60)
    assert bp.flush() is None
    bp.reset()
    with pytest.raises(ValueError):
        bp.add(10, 20)
    bp.reset()
    assert bp.add(95, 100) is None
    assert bp.flush() == (95, 100)


def test_loss_argument_errors_match_reference():
    from medicalseg_b200.models import losses as L
    with pytest.raises(TypeError):
        L.MixedLoss((L.DiceLoss(),), [1])
    with pytest.raises(ValueError):
        L.MixedLoss([L.DiceLoss()], [1, 2])
    with pytest.raises(RuntimeError):
        L.loss_computation([1, 2], None, {"types": [L.DiceLoss()], "coef": [1]})
    d = L.DiceLoss(sigmoid_norm=False, weight=[1.0, 2.0])  # both constructor options of dice_loss.py:36-43 exist
    assert d.sigmoid_norm is False and d.weight.tolist() == [1.0, 2.0]
    assert L.fused_head_plan({"types": [d], "coef": [1]}) is None  # non-default Dice: evaluate() takes the unfused path
    with pytest.raises(NotImplementedError):
        L.CrossEntropyLoss(data_format="NDHWC")


def test_transform_parameter_draws_match_the_oracle_on_cpu():
    """host side of medicalseg_b200.transforms (no device needed): same `random` seed -> same crop boxes as the
    restated reference (transform.py:240-279), rotation coefficients = scipy.ndimage.rotate's matrix / offset"""
    import random
    from medicalseg_b200 import transforms as T
    from oracle import transforms_oracle as to

    class Shape:  # get_params only looks at .shape
        def __init__(self, s):
            self.shape = s
    for seed in range(8):
        for shape in ((20, 24, 22), (128, 128, 128), (12, 512, 512)):
            random.seed(seed)
            ours = T.RandomResizedCrop3D(size=16, scale=[0.8, 1.2]).get_params(Shape(shape), [0.8, 1.2], (3. / 4., 4. / 3.))
            random.seed(seed)
            ref = to.RandomResizedCrop3D(size=16, scale=[0.8, 1.2]).box(shape)
            assert tuple(ours) == tuple(ref)
        random.seed(seed)
        a1 = T.RandomRotation3D(degrees=90).get_params((-90, 90))
        random.seed(seed)
        angle = random.uniform(-90, 90)
        plane = [[0, 1], [0, 2], [1, 2]][random.randint(0, 2)]
        assert a1 == (angle, plane)
    m, off = T.rotation_coefficients((7, 5), 90.0)
    assert m == (0.0, 1.0, -1.0, 0.0) and off == (3.0 - 2.0, 2.0 + 3.0)  # degree-exact quarter turn
    m, off = T.rotation_coefficients((10, 10), 30.0)
    c, s = np.cos(np.pi / 6), 0.5
    assert abs(m[0] - c) < 1e-15 and abs(m[1] - s) < 1e-15 and abs(m[2] + s) < 1e-15
    assert abs(off[0] - (4.5 - (c * 4.5 + s * 4.5))) < 1e-12 and abs(off[1] - (4.5 - (-s * 4.5 + c * 4.5))) < 1e-12
    with pytest.raises(ValueError):
        T.RandomRotation3D(degrees=-5)
    with pytest.raises(TypeError):
        T.Compose(transforms=(T.RandomFlip3D(),))


def test_fused_head_plan_covers_the_shipped_loss_configs_only():
    from medicalseg_b200.models import losses as L
    ce, dice = L.CrossEntropyLoss(), L.DiceLoss()
    plan = L.fused_head_plan({"types": [L.MixedLoss([ce, dice], [1, 2])], "coef": [0.5]})
    assert plan[0] is ce and plan[1] is dice and plan[2] == [(0, 0.5), (1, 1.0)]
    assert L.fused_head_plan({"types": [dice], "coef": [1]}) == (None, dice, [(1, 1)])
    assert L.fused_head_plan({"types": [ce], "coef": [3]}) == (ce, None, [(0, 3)])
    assert L.fused_head_plan({"types": [dice, dice], "coef": [1, 1]}) is None       # several logits (deep supervision)
    assert L.fused_head_plan({"types": [L.MixedLoss([dice, dice], [1, 1])], "coef": [1]}) is None
    assert L.fused_head_plan(None) is None
