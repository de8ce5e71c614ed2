"""DiceLoss / CrossEntropyLoss / MixedLoss with the reference signatures, running on ONE fused CUDA pass.

Reference: medicalseg/models/losses/dice_loss.py:23-102, cross_entropy_loss.py:23-87, loss_utils.py:18-40,
mixes_losses.py:22-60.  Behaviours kept on purpose:
  * Dice uses sigmoid(logits), the V-Net squared denominator, eps clip 1e-6, includes background, mean over classes;
    returns (loss, per_channel_dice) where per_channel_dice is a LazyHostArray: the D2H copy into pinned memory is
    queued on the stream at once, the host only waits for it when the values are first LOOKED AT (np.mean(dice),
    dice[0], arithmetic ...), so a training loop that logs every N iterations does not drain the GPU every step
    (the reference syncs in dice_loss.py:99 on every call).
  * CrossEntropyLoss(weight=None) computes class weights sum(1-p)/sum(p) from the FIRST logits it ever sees and
    caches them on the module (cross_entropy_loss.py:68-69).
  * MixedLoss returns ([coef_i * loss_i], per_channel_dice).
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from .. import ops


class LazyHostArray:
    """ndarray stand-in for a small device result: async D2H into pinned memory + event; materialises on first use.
    Implements the numpy array protocol, so np.mean(x), acc += x, x[i], len(x), float(x[i]) behave like the ndarray
    the reference returns (dice_loss.py:99-102)."""

    __array_priority__ = 100

    def __init__(self, dev: torch.Tensor):
        self._np = None
        self._dev = None
        if torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture nothing may touch the host: keep the (static) device tensor; the graph owner
            # (GraphedTrainStep) hands out a fresh LazyHostArray of it after every replay
            self._dev, self._pinned, self._event = dev, None, None
            return
        self._pinned = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        self._pinned.copy_(dev, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def numpy(self) -> np.ndarray:
        if self._np is None:
            if self._dev is not None:
                self._np = self._dev.detach().cpu().numpy()
            else:
                self._event.synchronize()
                self._np = self._pinned.numpy().copy()
                self._pinned = None
        return self._np

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getattr__(self, name):  # shape, dtype, mean, tolist, ...
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.numpy(), name)

    def __getitem__(self, i):
        return self.numpy()[i]

    def __len__(self):
        return len(self.numpy())

    def __iter__(self):
        return iter(self.numpy())

    def __repr__(self):
        return repr(self.numpy())

    def __add__(self, o): return self.numpy() + o
    def __radd__(self, o): return o + self.numpy()
    def __sub__(self, o): return self.numpy() - o
    def __rsub__(self, o): return o - self.numpy()
    def __mul__(self, o): return self.numpy() * o
    def __rmul__(self, o): return o * self.numpy()
    def __truediv__(self, o): return self.numpy() / o
    def __rtruediv__(self, o): return o / self.numpy()


class _DiceCEFunction(torch.autograd.Function):
    """result[0] = CE, result[1] = Dice loss, result[2:] = per-channel Dice; gradient flows to logits only."""

    @staticmethod
    def forward(ctx, logits, labels, class_w, ignore_index, dice_w=None, dice_softmax=False):
        n, c = logits.shape[:2]
        acc = ops.zero_(torch.empty(3 * c + 2, dtype=torch.float64, device=logits.device))
        result = torch.empty(2 + c, dtype=torch.float32, device=logits.device)
        ops.dice_ce_fwd(logits, labels, class_w, ignore_index, acc, dice_softmax)
        ops.dice_ce_finalize(acc, c, result, dice_w)
        ctx.save_for_backward(logits, labels, class_w, acc)
        ctx.ignore_index, ctx.dice_w, ctx.dice_softmax = ignore_index, dice_w, dice_softmax
        return result

    @staticmethod
    def backward(ctx, g):
        logits, labels, class_w, acc = ctx.saved_tensors
        dlogits = torch.empty_like(logits)
        coef = g[:2].contiguous().float()
        ops.dice_ce_bwd(logits, labels, class_w, acc, ctx.ignore_index, 1.0, 1.0, coef, dlogits, ctx.dice_w,
                        ctx.dice_softmax)
        return dlogits, None, None, None, None, None


def _prep(logits, labels):
    if not logits.is_cuda:
        raise RuntimeError("medicalseg_b200 losses need CUDA tensors (no CPU fallback)")
    if logits.dim() == 4:
        logits = logits.unsqueeze(0)
    if labels.dim() == 3:
        labels = labels.unsqueeze(0)
    assert "int" in str(labels.dtype), "The label should be int but got {}".format(labels.dtype)
    return logits.float().contiguous(), labels.to(torch.int32).contiguous()


def class_weights(logits: torch.Tensor) -> torch.Tensor:
    """models/losses/loss_utils.py:31-40 on the GPU: w_c = sum(1-softmax_c) / sum(softmax_c)."""
    n, c = logits.shape[:2]
    psum = ops.zero_(torch.empty(c, dtype=torch.float64, device=logits.device))
    w = torch.empty(c, dtype=torch.float32, device=logits.device)
    ops.class_weight_sums(logits, psum)
    ops.class_weight_finalize(psum, float(n * logits[0, 0].numel()), c, w)
    return w


class _FusedEval:
    """one fused forward shared by the CE and Dice objects of a MixedLoss.  The cache is keyed on the IDENTITY of the
    logits / labels / weight tensor objects (weak references: an id() or data_ptr() can be reused by the next step's
    tensors once the old ones are freed, which would hand out a result whose autograd graph is already consumed)."""
    refs = None
    key = None
    result = None


def _fused(logits, labels, class_w, ignore_index, dice_opts=(None, False)):
    """dice_opts = (per-class Dice weight tensor or None, softmax-normalised Dice) - DiceLoss's constructor options"""
    dice_w, dice_softmax = dice_opts
    key = (logits._version, ignore_index, None if dice_w is None else dice_w.data_ptr(), bool(dice_softmax))
    r = _FusedEval.refs
    if (r is not None and r[0]() is logits and r[1]() is labels and r[2]() is class_w and _FusedEval.key == key
            and _FusedEval.result is not None):
        return _FusedEval.result
    res = _DiceCEFunction.apply(logits, labels, class_w, ignore_index, dice_w, bool(dice_softmax))
    _FusedEval.refs = (weakref.ref(logits), weakref.ref(labels), weakref.ref(class_w))
    _FusedEval.key, _FusedEval.result = key, res
    return res


class DiceLoss:
    def __init__(self, sigmoid_norm=True, weight=None):
        # dice_loss.py:36-43: nn.Sigmoid() or nn.Softmax(axis=1); `weight` scales the per-class intersections (:64-65)
        self.sigmoid_norm, self.eps = bool(sigmoid_norm), 1e-5
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32).flatten()
        self._ones = None

    def _opts(self, device, c):
        if self.weight is not None:
            if self.weight.numel() != c:
                raise ValueError("DiceLoss: {} class weights for {} classes".format(self.weight.numel(), c))
            if self.weight.device != device:
                self.weight = self.weight.to(device)
        return (self.weight, not self.sigmoid_norm)

    def __call__(self, logits, labels):
        return self.forward(logits, labels)

    def forward(self, logits, labels, _class_w=None, _ignore_index=255):
        logits, labels = _prep(logits, labels)
        c = logits.shape[1]
        if _class_w is None:
            if self._ones is None or self._ones.numel() != c:
                self._ones = torch.ones(c, dtype=torch.float32, device=logits.device)
            _class_w = self._ones
        res = _fused(logits, labels, _class_w, _ignore_index, self._opts(logits.device, c))
        per_channel_dice = LazyHostArray(res[2:].detach())  # dice_loss.py:99 without draining the GPU here
        return res[1], per_channel_dice


class CrossEntropyLoss:
    def __init__(self, weight=None, ignore_index=255, data_format="NCDHW"):
        if data_format != "NCDHW":
            raise NotImplementedError("only data_format='NCDHW' is supported")
        self.ignore_index, self.EPS, self.data_format = ignore_index, 1e-8, data_format
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32)

    def __call__(self, logit, label):
        return self.forward(logit, label)

    def forward(self, logit, label, _dice_opts=(None, False)):
        logit, label = _prep(logit, label)
        if self.weight is None:
            self.weight = class_weights(logit.detach())  # cached forever (cross_entropy_loss.py:68-69)
        self.weight = self.weight.to(logit.device)
        if logit.shape[1] != len(self.weight):
            raise ValueError("The number of weights = {} must be the same as the number of classes = {}.".format(
                len(self.weight), logit.shape[1]))
        return _fused(logit, label, self.weight, self.ignore_index, _dice_opts)[0]


class MixedLoss:
    def __init__(self, losses, coef):
        if not isinstance(losses, list):
            raise TypeError("`losses` must be a list!")
        if not isinstance(coef, list):
            raise TypeError("`coef` must be a list!")
        if len(losses) != len(coef):
            raise ValueError("The length of `losses` should equal to `coef`, but they are {} and {}.".format(
                len(losses), len(coef)))
        self.losses, self.coef = losses, coef

    def __call__(self, logits, labels):
        return self.forward(logits, labels)

    def forward(self, logits, labels):
        logits, labels = _prep(logits, labels)
        loss_list, per_channel_dice = [], None
        ce = next((l for l in self.losses if type(l).__name__ == "CrossEntropyLoss"), None)
        dice = next((l for l in self.losses if type(l).__name__ == "DiceLoss"), None)
        # the CE and Dice objects share ONE fused pass: the CE call runs it with the Dice object's options
        dice_opts = dice._opts(logits.device, logits.shape[1]) if dice is not None else (None, False)
        for i, loss in enumerate(self.losses):
            if type(loss).__name__ == "DiceLoss":
                if ce is not None:  # share the CE object's class weights so both run in the same fused pass
                    if ce.weight is None:
                        ce.weight = class_weights(logits.detach())
                    out, per_channel_dice = loss.forward(logits, labels, ce.weight.to(logits.device), ce.ignore_index)
                else:
                    out, per_channel_dice = loss(logits, labels)
            elif loss is ce:
                out = loss.forward(logits, labels, dice_opts)
            else:
                out = loss(logits, labels)
            loss_list.append(out * self.coef[i])
        return loss_list, per_channel_dice


def check_logits_losses(logits_list, losses):
    if len(logits_list) != len(losses["types"]):
        raise RuntimeError("The length of logits_list should equal to the types of loss config: {} != {}.".format(
            len(logits_list), len(losses["types"])))


def loss_computation(logits_list, labels, losses, edges=None):
    """medicalseg/utils/loss_utils.py:25-52"""
    check_logits_losses(logits_list, losses)
    loss_list, per_channel_dice = [], None
    for i in range(len(logits_list)):
        logits, loss_i, coef_i = logits_list[i], losses["types"][i], losses["coef"][i]
        name = loss_i.__class__.__name__
        if name == "MixedLoss":
            mixed_loss_list, per_channel_dice = loss_i(logits, labels)
            for mixed_loss in mixed_loss_list:
                loss_list.append(coef_i * mixed_loss)
        elif name == "DiceLoss":
            loss, per_channel_dice = loss_i(logits, labels)
            loss_list.append(coef_i * loss)
        else:
            loss_list.append(coef_i * loss_i(logits, labels))
    return loss_list, per_channel_dice


# ---- evaluation fast path: losses + argmax fused behind the model's 1x1x1 head ---------------------------------
def fused_head_plan(losses):
    """(ce, dice, terms) for the loss configurations the fused evaluation head covers - one logits tensor scored by a
    DiceLoss, a CrossEntropyLoss or a MixedLoss of those (what core/val.py:95 builds from the shipped configs) - else
    None.  terms = [(kind, coefficient)] in loss_computation's output order."""
    if losses is None or len(losses["types"]) != 1:
        return None
    obj, coef = losses["types"][0], losses["coef"][0]
    name = type(obj).__name__
    if name == "MixedLoss":
        parts = [(type(l).__name__, l, c) for l, c in zip(obj.losses, obj.coef)]
    else:
        parts = [(name, obj, 1)]
    ce = dice = None
    terms = []
    for kind, l, c in parts:
        if kind == "CrossEntropyLoss" and ce is None:
            ce = l
            terms.append((0, coef * c))
        elif kind == "DiceLoss" and dice is None:
            if not l.sigmoid_norm or l.weight is not None:
                return None  # the fused head computes the default sigmoid Dice: other options take the unfused path
            dice = l
            terms.append((1, coef * c))
        else:
            return None
    return ce, dice, terms


def fused_head_losses(ao, w2, b2, c, dims, labels, losses, plan):
    """runs ops.eval_head on the activated out_tr features `ao` (B8); see VNet.predict_with_losses"""
    n, dev = ao.n, ao.buf.device
    pred = torch.empty((n, 1, *dims), dtype=torch.int32, device=dev)
    if labels is None:
        ops.eval_head(ao, w2, b2, None, None, c, 255, pred=pred)
        return pred, None, None
    ce, dice, terms = plan
    if labels.dim() == 3:
        labels = labels.unsqueeze(0)
    assert "int" in str(labels.dtype), "The label should be int but got {}".format(labels.dtype)
    labels = labels.to(torch.int32).contiguous()
    ignore_index = ce.ignore_index if ce is not None else 255
    if ce is not None:
        if ce.weight is None:  # first logits ever seen define the class weights (cross_entropy_loss.py:68-69)
            psum = ops.zero_(torch.empty(c, dtype=torch.float64, device=dev))
            ops.eval_head(ao, w2, b2, None, None, c, ignore_index, psum=psum)
            ce.weight = torch.empty(c, dtype=torch.float32, device=dev)
            ops.class_weight_finalize(psum, float(n * ao.s), c, ce.weight)
        ce.weight = ce.weight.to(dev)
        if c != len(ce.weight):
            raise ValueError("The number of weights = {} must be the same as the number of classes = {}.".format(
                len(ce.weight), c))
        class_w = ce.weight
    else:
        if dice._ones is None or dice._ones.numel() != c or dice._ones.device != dev:
            dice._ones = torch.ones(c, dtype=torch.float32, device=dev)
        class_w = dice._ones
    acc = ops.zero_(torch.empty(3 * c + 2, dtype=torch.float64, device=dev))
    result = torch.empty(2 + c, dtype=torch.float32, device=dev)
    ops.eval_head(ao, w2, b2, labels, class_w, c, ignore_index, pred=pred, acc=acc)
    ops.dice_ce_finalize(acc, c, result)
    loss_list = [result[kind] * coef for kind, coef in terms]
    per_channel_dice = LazyHostArray(result[2:]) if dice is not None else None
    return pred, loss_list, per_channel_dice
