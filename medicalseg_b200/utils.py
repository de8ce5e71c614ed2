"""Checkpoint helpers (reference: medicalseg/utils/utils.py:76-135).  `.pdparams` files are pickled dicts of
numpy arrays keyed by the Paddle parameter names this package keeps, so they load without PaddlePaddle."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch


class _NumpyOnlyUnpickler(pickle.Unpickler):
    """Unpickler for Paddle-format checkpoints (pickled dicts of numpy arrays): only the constructors numpy's own
    array / dtype / scalar reduction uses and plain containers are allowed, so a crafted file cannot run code."""

    _ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"), ("collections", "OrderedDict"),
        ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
        ("_codecs", "encode"),  # protocol-2 pickles carry the array bytes as a latin-1 string re-encoded on load
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError("checkpoint refers to %s.%s: only numpy arrays and plain containers are accepted"
                                     % (module, name))


def _load_any(path):
    """reads a checkpoint file: a torch tensor pickle (weights_only) or a Paddle-style pickle of numpy arrays"""
    if path.startswith("http://") or path.startswith("https://"):
        raise RuntimeError("downloading pretrained weights is not supported offline; pass a local file: %s" % path)
    if os.path.isdir(path):
        path = os.path.join(path, "model.pdparams")
    if not os.path.exists(path):
        raise ValueError("The pretrained model directory is not Found: {}".format(path))
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except (pickle.UnpicklingError, RuntimeError, EOFError, ValueError, KeyError, AttributeError):
        # not a torch zip / a pickle with numpy objects: the Paddle container (protocol-2 pickle of numpy arrays)
        with open(path, "rb") as fh:
            return _NumpyOnlyUnpickler(fh, encoding="latin1").load()


def load_entire_model(model, pretrained):
    """utils.py:76-82 — load every matching-shape key, warn about the rest (utils.py:84-104)."""
    if pretrained is None:
        return
    sd = _load_any(pretrained)
    own = model.state_dict()
    ok = {}
    for k, v in sd.items():
        if k == "StructuredToParameterName@@":  # Paddle's name table, not a parameter
            continue
        if k not in own:
            print("[WARNING] {} is not in pretrained model".format(k))
            continue
        arr = np.asarray(v) if not torch.is_tensor(v) else v
        if tuple(arr.shape) != tuple(own[k].shape):
            print("[WARNING] [SKIP] Shape of pretrained params {} doesn't match.(Pretrained: {}, Actual: {})".format(
                k, tuple(arr.shape), tuple(own[k].shape)))
            continue
        ok[k] = arr
    model.set_state_dict(ok, strict=False)
    print("[INFO] There are {}/{} variables loaded into {}.".format(len(ok), len(own), model.__class__.__name__))


def _to_numpy_tree(v):
    if torch.is_tensor(v):
        return v.detach().cpu().numpy()
    if isinstance(v, dict):
        return {k: _to_numpy_tree(x) for k, x in v.items()}
    return v


def save_checkpoint(model, optimizer, save_dir):
    """core/train.py:230-236 layout: <dir>/model.pdparams + model.pdopt, both in the container `paddle.save` writes
    (protocol-2 pickles of numpy arrays, parameter names = the reference's), so the reference's `paddle.load` +
    `set_dict` reads a checkpoint trained here and `resume` / `pretrained=` read the reference's."""
    os.makedirs(save_dir, exist_ok=True)
    export_pdparams(model, os.path.join(save_dir, "model.pdparams"))
    if optimizer is not None:
        with open(os.path.join(save_dir, "model.pdopt"), "wb") as fh:
            pickle.dump(_to_numpy_tree(optimizer.state_dict()), fh, protocol=2)


def export_pdparams(model, path):
    """Writes the parameters in the container PaddlePaddle 2.x itself uses for `paddle.save(layer.state_dict())`: a
    protocol-2 pickle of {structured name: numpy array} (+ the `StructuredToParameterName@@` name table), so weights
    trained here can be taken back to the reference (`paddle.load` + `set_dict`, utils/utils.py:84-104).  The same
    container is what `load_entire_model` / `VNet(pretrained=...)` read when handed a reference checkpoint."""
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    sd["StructuredToParameterName@@"] = {k: k for k in sd}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fh:
        pickle.dump(sd, fh, protocol=2)


def resume(model, optimizer, resume_model):
    """utils.py:115-135 — returns the iteration parsed from the directory suffix iter_N."""
    if resume_model is None:
        return 0
    resume_model = os.path.normpath(resume_model)
    if not os.path.exists(resume_model):
        raise ValueError("Directory of the model needed to resume is not Found: {}".format(resume_model))
    sd = _load_any(os.path.join(resume_model, "model.pdparams"))
    sd.pop("StructuredToParameterName@@", None)
    model.set_state_dict(sd)
    if optimizer is not None:
        optimizer.set_state_dict(_load_any(os.path.join(resume_model, "model.pdopt")))
    return int(resume_model.split("_")[-1])


class DevicePrefetcher:
    """Double-buffered host -> device input staging on a copy stream (the role of the reference's DataLoader worker
    prefetch, core/train.py:100-107, moved to the device side): `stage(images, labels)` starts the asynchronous copy
    of the NEXT batch (pinned host tensors) while the current step runs; `get()` makes the compute stream wait for it
    and hands out the device tensors.  Two buffer pairs alternate, so a batch is never overwritten while a step that
    was launched on it may still be running - provided the caller synchronises (e.g. reads the loss) once per step."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.events = [None, None]
        self.next = 0
        self.ready = None

    def stage(self, images: torch.Tensor, labels: torch.Tensor):
        i = self.next
        if self.bufs[i] is None or self.bufs[i][0].shape != images.shape or self.bufs[i][1].shape != labels.shape:
            self.bufs[i] = (torch.empty(images.shape, dtype=images.dtype, device=self.device),
                            torch.empty(labels.shape, dtype=labels.dtype, device=self.device))
        self.stream.wait_stream(torch.cuda.current_stream(self.device))  # earlier consumers of this buffer are queued
        with torch.cuda.stream(self.stream):
            self.bufs[i][0].copy_(images, non_blocking=True)
            self.bufs[i][1].copy_(labels, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.events[i] = ev
        self.ready = i
        self.next = 1 - i

    def get(self):
        i = self.ready
        if i is None:
            raise RuntimeError("DevicePrefetcher.get() without a staged batch")
        torch.cuda.current_stream(self.device).wait_event(self.events[i])
        self.ready = None
        return self.bufs[i]
