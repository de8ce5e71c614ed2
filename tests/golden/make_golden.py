"""Generates the committed golden fixtures.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

* preprocess_*.npz : outputs of the reference's OWN tools/preprocess_utils/{values,geometry}.py
  (NumPy/SciPy branch), imported by file path with tools.preprocess_utils.global_var stubbed
  (USE_GPU=False) — the package __init__ itself needs nibabel/SimpleITK, absent here.
* vnet_oracle_*.npz : outputs of oracle/vnet_oracle.py on seeded inputs (the reference VNet needs
  PaddlePaddle, not installable offline -> these pin the ORACLE against drift, not the reference).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def load_reference_preprocess():
    gv = types.ModuleType("tools.preprocess_utils.global_var")
    store = {"USE_GPU": False}
    gv.get_value = lambda k, d=None: store.get(k, d)
    gv.set_value = lambda k, v: store.__setitem__(k, v)
    for name in ("tools", "tools.preprocess_utils"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["tools.preprocess_utils.global_var"] = gv
    sys.modules["tools.preprocess_utils"].global_var = gv
    mods = {}
    for fn in ("values", "geometry"):
        spec = importlib.util.spec_from_file_location(
            "ref_" + fn, os.path.join(REF, "tools/preprocess_utils", fn + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[fn] = mod
    return mods["values"], mods["geometry"]


def ct_volume(shape, seed):
    rng = np.random.default_rng(seed)
    v = rng.uniform(-2000, 2000, size=shape).astype(np.float32)
    nan_mask = rng.random(shape) < 0.001
    v[nan_mask] = np.nan
    return v


def make_preprocess():
    values, geometry = load_reference_preprocess()
    out = {}
    cases = [
        ("iso", (40, 40, 40), (16, 16, 16)),
        ("aniso", (33, 47, 12), (16, 24, 12)),
        ("up", (9, 10, 11), (20, 17, 23)),
        ("one", (8, 8, 8), (1, 4, 8)),
    ]
    for name, shp, new in cases:
        vol = ct_volume(shp, seed=len(name))
        out[f"{name}_in"] = vol
        hu = values.HUnorm(vol.copy())
        out[f"{name}_hunorm"] = hu.astype(np.float32)
        r1, sp = geometry.resample(hu.astype(np.float32), spacing=(1.0, 0.7, 0.7), new_shape=list(new), order=1)
        out[f"{name}_resample1"] = r1
        out[f"{name}_spacing"] = np.asarray(sp, dtype=np.float64)
        lab = (np.random.default_rng(7).integers(0, 4, size=shp)).astype(np.int32)
        out[f"{name}_label"] = lab
        r0, _ = geometry.resample(lab, new_shape=list(new), order=0)
        out[f"{name}_resample0"] = r0
        out[f"{name}_norm_minmax"] = values.normalize(np.nan_to_num(vol.copy(), nan=0.0)).astype(np.float32)
        out[f"{name}_norm_fixed"] = values.normalize(np.nan_to_num(vol.copy(), nan=0.0), 0, 2650).astype(np.float32)
    lab = np.random.default_rng(3).integers(0, 6, size=(6, 7, 8)).astype(np.int32)
    out["remap_in"] = lab
    out["remap_out"] = values.label_remap(lab.copy(), {1: 2, 2: 3, 5: 0})
    # resample with new_shape=None (spacing driven)
    vol = ct_volume((20, 24, 28), seed=11)
    hu = values.HUnorm(vol.copy()).astype(np.float32)
    r, sp = geometry.resample(hu, spacing=(2.0, 1.5, 0.5), new_spacing=[1.0, 1.0, 1.0], order=1)
    out["spacing_in"] = hu
    out["spacing_out"] = r
    np.savez_compressed(os.path.join(HERE, "preprocess_ref.npz"), **out)
    print("preprocess_ref.npz:", {k: v.shape for k, v in out.items()})


def make_vnet():
    from oracle import vnet_oracle as vo

    out = {}
    configs = {
        "iso2": dict(num_classes=2, shape=(16, 16, 16)),
        "iso3": dict(num_classes=3, shape=(16, 16, 16)),
        "mri20": dict(num_classes=20, shape=(32, 32, 12),
                      kernel_size=[[2, 2, 4], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
                      stride_size=[[2, 2, 1], [2, 2, 1], [2, 2, 2], [2, 2, 2]]),
    }
    for name, cfg in configs.items():
        torch.manual_seed(0)
        kw = {k: v for k, v in cfg.items() if k in ("kernel_size", "stride_size")}
        model = vo.VNetOracle(num_classes=cfg["num_classes"], **kw)
        img, lab = vo.synthetic_batch(2, cfg["shape"], cfg["num_classes"], seed=0)
        # eval-mode forward + loss (mirrors the reference's commented alignment harness, vnet.py:299,361)
        model.eval()
        with torch.no_grad():
            logits = model(img)[0]
        losses = vo.default_losses()
        ll, dice = vo.loss_computation([logits], lab, losses)
        out[f"{name}_eval_logits_sample"] = logits.flatten()[:: max(1, logits.numel() // 4096)].numpy()
        out[f"{name}_eval_logits_absmean"] = np.float64(logits.abs().mean().item())
        out[f"{name}_eval_losses"] = np.array([float(l) for l in ll])
        out[f"{name}_eval_dice"] = np.asarray(dice)
        # 3 train steps with explicit masks
        model.train()
        losses = vo.default_losses()
        opt = vo.Momentum(vo.PolynomialDecay(0.001, 15000), list(model.parameters()), 0.9, 1e-4)
        rec = []
        for step in range(3):
            masks = vo.make_dropout_masks(2, seed=0, step=step)
            loss, ll, dice = vo.train_step(model, losses, opt, img, lab, masks)
            rec.append([loss] + ll + list(np.asarray(dice, dtype=np.float64)))
        out[f"{name}_train_record"] = np.asarray(rec)
        out[f"{name}_param_checksum"] = np.float64(sum(p.double().abs().sum().item() for p in model.parameters()))
        print(name, "eval losses", out[f"{name}_eval_losses"], "train", np.asarray(rec)[:, 0])
    np.savez_compressed(os.path.join(HERE, "vnet_oracle.npz"), **out)


if __name__ == "__main__":
    make_preprocess()
    make_vnet()
