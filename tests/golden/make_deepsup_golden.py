"""Pins oracle.vnet_oracle.VNetDeepSupOracle against drift (tests/golden/vnet_deepsup_oracle.npz).  The reference
VNetDeepSup needs PaddlePaddle (not installable offline), so - as for vnet_oracle.npz - these vectors pin the ORACLE, not
the reference; the pieces it adds over the VNet oracle (3x3x3 heads, half-pixel trilinear resize) are cross-checked
against independent formulas in tests/test_oracle_vnet.py.

    python tests/golden/make_deepsup_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from oracle import vnet_oracle as vo
    torch.manual_seed(0)
    model = vo.VNetDeepSupOracle(num_classes=3)
    img, lab = vo.synthetic_batch(2, (16, 16, 16), 3, seed=0)
    model.train()
    masks = vo.make_dropout_masks(2, seed=0, step=0)
    logits = model(img, masks)
    ll, dice = vo.loss_computation(logits, lab, vo.deepsup_losses())
    out = {"losses": np.array([float(l) for l in ll]), "dice": np.asarray(dice, dtype=np.float64)}
    for i, t in enumerate(logits):
        out["logits_%d_sample" % i] = t.detach().flatten()[:: max(1, t.numel() // 2048)].numpy()
    sum(ll).backward()
    for name in ("out_tr64.weight", "out_tr256.bias", "up_tr256.ops.1.conv1.weight", "in_tr.conv1.weight"):
        g = dict(model.named_parameters())[name].grad
        out["grad_norm_" + name] = np.float64(g.double().norm().item())
    np.savez_compressed(os.path.join(HERE, "vnet_deepsup_oracle.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
