"""GPU tests of the round-2 boundary work: multi-channel input (vnet.py:74-79), inference() with the Resize3D reverse
transform (core/infer.py:44-94), kernel-drawn Dropout3D masks, the background .npy loader (core/train.py:90-95) and the
data-parallel CUDA-graph step (needs >= 2 GPUs: run with `gpurun --gpus 2`)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dtype,in_ch", [("f32", 2), ("bf16", 2), ("bf16", 4), ("f32", 16)])
def test_vnet_multi_channel_input_matches_oracle(dtype, in_ch):
    """in_channels in {2, 4, 8, 16}: x is tiled 16/Cin times onto the conv output (vnet.py:74-79)"""
    from oracle import vnet_oracle as vo
    from medicalseg_b200.models import VNet, losses as L
    torch.manual_seed(0)
    om = vo.VNetOracle(num_classes=2, in_channels=in_ch)
    om.train()
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, in_ch, 32, 32, 32, generator=g)
    _, lab = vo.synthetic_batch(2, (32, 32, 32), 2, seed=0)
    masks = vo.make_dropout_masks(2, seed=0)
    ologits = om(img, masks)[0]
    oll, odice = vo.loss_computation([ologits], lab, vo.default_losses())
    sum(oll).backward()
    m = VNet(num_classes=2, in_channels=in_ch, compute_dtype=dtype)
    m.set_state_dict(om.state_dict())
    m.train()
    m.set_dropout_masks(masks)
    ours = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    logits = m(img.cuda())[0]
    ll, dice = L.loss_computation([logits], lab.cuda(), ours)
    sum(ll).backward()
    lg, ol = logits.detach().cpu(), ologits.detach()
    gw = m.store.grad_view("in_tr.conv1.weight").cpu()
    ogw = om.in_tr.conv1.weight.grad
    cos = float((gw * ogw).sum() / (gw.norm() * ogw.norm()))
    if dtype == "f32":
        assert float((lg - ol).abs().max()) <= 1e-4 * float(ol.abs().max())
        assert float(np.abs(np.asarray(dice) - np.asarray(odice)).max()) <= 1e-5
        assert cos >= 0.9999, cos
    else:
        rms = float(torch.sqrt(((lg - ol) ** 2).mean()) / torch.sqrt((ol ** 2).mean()))
        assert rms <= 2e-2, rms
        assert float(np.abs(np.asarray(dice) - np.asarray(odice)).max()) <= 1e-3
        assert cos >= 0.97, cos


def test_inference_argmax_and_resize3d_reverse_transform():
    """core/infer.py:62-94: pred = argmax(logits); when the validation transforms hold a Resize3D the logits are first
    resized back to `ori_shape` (linear interpolation, F.interpolate defaults) - checked against torch's trilinear"""
    from medicalseg_b200.core import inference, reverse_transform
    from medicalseg_b200.models import VNet
    from medicalseg_b200.transforms import Resize3D
    m = VNet(num_classes=3, compute_dtype="bf16")
    m.eval()
    x = torch.rand(1, 1, 32, 32, 32, device="cuda")
    with torch.no_grad():
        pred, logit = inference(m, x)
        assert pred.dtype == torch.int32 and tuple(pred.shape) == (1, 1, 32, 32, 32)
        assert torch.equal(pred.long(), logit.argmax(1, keepdim=True))
        tf = [Resize3D((32, 32, 32))]
        back = reverse_transform(logit, (40, 48, 36), tf)  # the same logits through our kernel and through torch
        ref = torch.nn.functional.interpolate(logit, size=(40, 48, 36), mode="trilinear", align_corners=False)
        assert float((back - ref).abs().max()) <= 2e-5 * float(ref.abs().max() + 1)
        pred2, logit2 = inference(m, x, ori_shape=(40, 48, 36), transforms=tf)
        assert tuple(logit2.shape) == (1, 3, 40, 48, 36) and tuple(pred2.shape) == (1, 1, 40, 48, 36)
        assert torch.equal(pred2.long(), logit2.argmax(1, keepdim=True))
        with pytest.raises(ValueError):
            inference(m, x, ori_shape=(40, 48, 36), transforms=[])  # nothing explains the shape difference


def test_dropout_masks_are_kernel_drawn_and_advance_per_step():
    from medicalseg_b200 import ops
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = torch.empty(4096, device="cuda")
    b = torch.empty(4096, device="cuda")
    ops.dropout_masks(1234, step, a, 0.5)
    ops.dropout_masks(1234, step, b, 0.5)
    assert int(step.item()) == 2
    assert set(a.unique().tolist()) == {0.0, 2.0}
    assert 0.9 < float(a.mean()) < 1.1 and not torch.equal(a, b)
    step.zero_()
    c = torch.empty(4096, device="cuda")
    ops.dropout_masks(1234, step, c, 0.5)
    assert torch.equal(a, c)  # same (seed, step) -> same masks
    # the model draws its own masks when none are given, and two steps differ
    from medicalseg_b200.models import VNet
    m = VNet(num_classes=2, compute_dtype="bf16", seed=0)
    m.train()
    x = torch.rand(2, 1, 32, 32, 32, device="cuda")
    with torch.no_grad():
        l1 = m(x)[0].clone()
        d1 = {k: v.clone() for k, v in m._drawn.items()}
        l2 = m(x)[0]
    assert set(d1) == {s for s, _ in VNet._DROPOUT_SITES} and tuple(d1["up_tr128.skip"].shape) == (2, 64)
    assert not torch.equal(l1, l2) and int(m._dropout_step.item()) == 2


def _write_npy_set(root, n, shape):
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "labels"), exist_ok=True)
    lines = []
    for i in range(n):
        np.save(os.path.join(root, "images", "%d.npy" % i), rng.random(shape, dtype=np.float32) * 255)
        np.save(os.path.join(root, "labels", "%d.npy" % i), (rng.random(shape) > 0.5).astype(np.int32))
        lines.append("images/%d.npy labels/%d.npy" % (i, i))
    for split in ("train", "val"):
        with open(os.path.join(root, "%s_list.txt" % split), "w") as fh:
            fh.write("\n".join(lines) + "\n")


def test_background_loader_keeps_reader_cost_near_zero_on_disk_npy(tmp_path):
    """DataLoader(num_workers) semantics (core/train.py:90-95): with workers the train loop's reader_cost is the time it
    blocks on the prefetch queue - < 1 ms per iteration on an on-disk .npy set at the 128^3 benchmark shape."""
    from medicalseg_b200.core import train
    from medicalseg_b200.datasets import BatchLoader, DistributedBatchSampler, NpyVolumeDataset
    from medicalseg_b200.models import VNet, losses as L
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    _write_npy_set(str(tmp_path / "ds"), 6, (128, 128, 128))
    ds = NpyVolumeDataset(str(tmp_path / "ds"), num_classes=2, mode="train")
    # loader contract: ordered, complete, device tensors, same content as the synchronous path
    sampler = DistributedBatchSampler(len(ds), 2, shuffle=False)
    sync = BatchLoader(ds, iter(sampler.epoch()), "cuda", num_workers=0)
    bg = BatchLoader(ds, iter(sampler.epoch()), "cuda", num_workers=3)
    n = 0
    for (a_im, a_lab), (b_im, b_lab) in zip(sync, bg):
        assert b_im.is_cuda and tuple(b_im.shape) == (2, 1, 128, 128, 128) and b_lab.dtype == torch.int32
        assert torch.equal(a_im, b_im) and torch.equal(a_lab, b_lab)
        n += 1
    assert n == len(sampler)
    bg.close()
    # the train loop's own log line
    m = VNet(num_classes=2, compute_dtype="bf16", seed=0)
    losses = {"types": [L.MixedLoss([L.CrossEntropyLoss(), L.DiceLoss()], [1, 1])], "coef": [1]}
    opt = Momentum(PolynomialDecay(0.001, 100), m.parameters(), 0.9, 1e-4)
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        train(m, ds, optimizer=opt, save_dir=str(tmp_path / "out"), iters=40, batch_size=2, save_interval=1000,
              log_iters=10, num_workers=3, losses=losses, to_static_training=True)
    lines = [l for l in buf.getvalue().splitlines() if "[TRAIN]" in l]
    assert len(lines) >= 3, buf.getvalue()
    rc = [float(re.search(r"reader_cost: ([0-9.]+)", l).group(1)) for l in lines]
    bc = [float(re.search(r"batch_cost: ([0-9.]+)", l).group(1)) for l in lines]
    print("reader_cost per log window:", rc, "batch_cost:", bc)
    assert min(rc[1:]) < 1e-3, (rc, bc)  # steady state (the first window holds the graph capture)


def test_train_checkpoints_use_the_paddle_container_and_resume(tmp_path):
    """model.pdparams / model.pdopt are protocol-2 pickles of numpy arrays under the reference's parameter names
    (core/train.py:230-236, utils/utils.py:115-135): readable without torch, and `resume` restores the step."""
    import pickle
    from medicalseg_b200.models import VNet
    from medicalseg_b200.optimizer import Momentum, PolynomialDecay
    from medicalseg_b200.utils import resume, save_checkpoint
    m = VNet(num_classes=2, compute_dtype="bf16", seed=1)
    opt = Momentum(PolynomialDecay(0.01, 100), m.parameters(), 0.9, 1e-4)
    opt.velocity.normal_()
    opt._learning_rate.last_epoch = 7
    d = str(tmp_path / "iter_7")
    save_checkpoint(m, opt, d)
    with open(os.path.join(d, "model.pdparams"), "rb") as fh:
        raw = pickle.load(fh)
    assert isinstance(raw["in_tr.conv1.weight"], np.ndarray) and raw["in_tr.conv1.weight"].shape == (16, 1, 5, 5, 5)
    assert "StructuredToParameterName@@" in raw and raw["up_tr256.ops.0.conv1.weight"].shape == (256, 256, 5, 5, 5)
    m2 = VNet(num_classes=2, compute_dtype="bf16", seed=2)
    opt2 = Momentum(PolynomialDecay(0.01, 100), m2.parameters(), 0.9, 1e-4)
    assert resume(m2, opt2, d) == 7
    for name in m.store.slots:
        assert torch.equal(m.store.view(name), m2.store.view(name)), name
    for name, slot in m.store.slots.items():  # (the flat buffers also hold alignment padding that is not saved)
        if not slot.is_buffer:
            assert torch.equal(m.store.view_of(opt.velocity, name), m2.store.view_of(opt2.velocity, name)), name
    assert opt2._learning_rate.last_epoch == 7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_graph_step_matches_single_process_batch4():
    """2 ranks x batch 2 with SyncBatchNorm (the reference's semantics at world > 1, cvlibs/config.py:322) == one process
    with batch 4: logits, summed gradients, running statistics; and the CAPTURED data-parallel step (NCCL all-reduces
    inside the CUDA graph) == the eager data-parallel step.  Runs tools/ddp_check.py under torchrun."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools/ddp_check.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "DDP_CHECK_OK" in r.stdout
