"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — torch-CPU restatement of the reference VNet path.

PARITY UNPINNED: the reference ships no golden vectors for this path and PaddlePaddle cannot be
installed offline, so this file restates the reference call sites one by one and encodes Paddle's
documented layer defaults (listed below).  Every function cites the reference file:line it follows
(paths relative to /root/reference).

Paddle defaults encoded here (third-party behaviour, not visible in /root/reference):
  * nn.Conv3D weight ~ Normal(0, sqrt(2 / (Cin*kD*kH*kW))), bias 0; weight layout [Cout,Cin,kD,kH,kW]
  * nn.Conv3DTranspose weight ~ XavierUniform, bias 0; weight layout [Cin,Cout,kD,kH,kW]
  * nn.BatchNorm3D(momentum=0.9, epsilon=1e-5): running = 0.9*running + 0.1*batch, the running
    variance is updated with the BIASED batch variance; buffers are named _mean / _variance
  * nn.PReLU(num_parameters=C, init=0.25); parameter named _weight
  * nn.Dropout3D(p=0.5): whole-channel mask [N,C,1,1,1], kept channels scaled by 1/(1-p)=2 (train only)
  * optimizer.Momentum(momentum, weight_decay=float): g += wd*p; v = mu*v + g; p -= lr*v (no Nesterov)
  * lr.PolynomialDecay(cycle=False): lr = (lr0-end)*(1-min(t,T)/T)**power + end
  * F.cross_entropy(weight=w, reduction='mean', ignore_index): sum_i w[y_i]*l_i / sum_i w[y_i]
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

DROPOUT_SITES = (
    # (site name, channel count) in execution order — vnet.py:103,108,144-145,149-150 via :201-232
    ("down_tr128", 128),
    ("down_tr256", 256),
    ("up_tr256.x", 256),
    ("up_tr256.skip", 128),
    ("up_tr128.x", 256),
    ("up_tr128.skip", 64),
)


def make_dropout_masks(batch: int, seed: int, step: int = 0, p: float = 0.5) -> Dict[str, torch.Tensor]:
    """Explicit Dropout3D masks [N, C] holding 0 or 1/(1-p) (Paddle upscale_in_train).

    Paddle's RNG stream cannot be reproduced, so masks are an explicit input on both sides of every
    parity test (SURVEY.md §7 'Dropout parity')."""
    g = torch.Generator().manual_seed(seed * 1000003 + step)
    out = {}
    for name, c in DROPOUT_SITES:
        keep = (torch.rand(batch, c, generator=g) >= p).to(torch.float32)
        out[name] = keep / (1.0 - p)
    return out


class PaddleBatchNorm3D(nn.Module):
    """nn.BatchNorm3D as used at vnet.py:38,70,100,139,167 (Paddle semantics, see module docstring)."""

    def __init__(self, c: int, momentum: float = 0.9, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("_mean", torch.zeros(c))
        self.register_buffer("_variance", torch.ones(c))
        self.momentum = momentum
        self.eps = eps

    def forward(self, x):
        shape = (1, -1, 1, 1, 1)
        if self.training:
            mean = x.mean(dim=(0, 2, 3, 4))
            var = x.var(dim=(0, 2, 3, 4), unbiased=False)
            with torch.no_grad():
                self._mean.mul_(self.momentum).add_((1 - self.momentum) * mean.detach())
                self._variance.mul_(self.momentum).add_((1 - self.momentum) * var.detach())
        else:
            mean, var = self._mean, self._variance
        xhat = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + self.eps)
        return xhat * self.weight.view(shape) + self.bias.view(shape)


class PaddlePReLU(nn.Module):
    """nn.PReLU(C): max(0,x) + a_c*min(0,x); parameter `_weight` (vnet.py:35,72,101-102,140-141,170)."""

    def __init__(self, c: int, init: float = 0.25):
        super().__init__()
        self._weight = nn.Parameter(torch.full((c,), init))

    def forward(self, x):
        a = self._weight.view(1, -1, 1, 1, 1)
        return torch.clamp(x, min=0) + a * torch.clamp(x, max=0)


def _paddle_conv_init(conv: nn.Conv3d):
    fan = conv.in_channels * int(np.prod(conv.kernel_size))
    nn.init.normal_(conv.weight, 0.0, math.sqrt(2.0 / fan))
    nn.init.zeros_(conv.bias)


def _paddle_convT_init(conv: nn.ConvTranspose3d):
    nn.init.xavier_uniform_(conv.weight)
    nn.init.zeros_(conv.bias)


class LUConv(nn.Module):  # vnet.py:32-43
    def __init__(self, nchan):
        super().__init__()
        self.relu1 = PaddlePReLU(nchan)
        self.conv1 = nn.Conv3d(nchan, nchan, kernel_size=5, padding=2)
        self.bn1 = PaddleBatchNorm3D(nchan)
        _paddle_conv_init(self.conv1)

    def forward(self, x):
        return self.relu1(self.bn1(self.conv1(x)))


class InputTransition(nn.Module):  # vnet.py:57-79
    def __init__(self, in_channels):
        super().__init__()
        self.num_features = 16
        self.in_channels = in_channels
        self.conv1 = nn.Conv3d(in_channels, 16, kernel_size=5, padding=2)
        self.bn1 = PaddleBatchNorm3D(16)
        self.relu1 = PaddlePReLU(16)
        _paddle_conv_init(self.conv1)

    def forward(self, x):
        out = self.bn1(self.conv1(x))
        rep = int(self.num_features / self.in_channels)
        x_tile = x.repeat(1, rep, 1, 1, 1)  # paddle Tensor.tile (vnet.py:78)
        return self.relu1(out + x_tile)


class DownTransition(nn.Module):  # vnet.py:82-113
    def __init__(self, in_ch, n_convs, dropout=False, stride=(2, 2, 2), kernel=(2, 2, 2)):
        super().__init__()
        out_ch = 2 * in_ch
        self.if_dropout = dropout
        self.down_conv = nn.Conv3d(in_ch, out_ch, kernel_size=tuple(kernel), stride=tuple(stride))
        self.bn1 = PaddleBatchNorm3D(out_ch)
        self.relu1 = PaddlePReLU(out_ch)
        self.relu2 = PaddlePReLU(out_ch)
        self.ops = nn.Sequential(*[LUConv(out_ch) for _ in range(n_convs)])
        _paddle_conv_init(self.down_conv)

    def forward(self, x, mask=None):
        down = self.relu1(self.bn1(self.down_conv(x)))
        out = down
        if self.if_dropout and self.training:
            out = down * mask.view(mask.shape[0], -1, 1, 1, 1).to(down.dtype)
        out = self.ops(out)
        return self.relu2(out + down)


class UpTransition(nn.Module):  # vnet.py:116-156
    def __init__(self, in_ch, out_ch, n_convs, dropout=False, dropout2=False, stride=(2, 2, 2), kernel=(2, 2, 2)):
        super().__init__()
        self.up_conv = nn.ConvTranspose3d(in_ch, out_ch // 2, kernel_size=tuple(kernel), stride=tuple(stride))
        self.bn1 = PaddleBatchNorm3D(out_ch // 2)
        self.relu1 = PaddlePReLU(out_ch // 2)
        self.relu2 = PaddlePReLU(out_ch)
        self.if_dropout = dropout
        self.if_dropout2 = dropout2
        self.ops = nn.Sequential(*[LUConv(out_ch) for _ in range(n_convs)])
        _paddle_convT_init(self.up_conv)

    def forward(self, x, skipx, mask_x=None, mask_skip=None):
        out = x
        if self.if_dropout and self.training:
            out = x * mask_x.view(mask_x.shape[0], -1, 1, 1, 1).to(x.dtype)
        if self.if_dropout2 and self.training:
            skipx = skipx * mask_skip.view(mask_skip.shape[0], -1, 1, 1, 1).to(x.dtype)
        out = self.relu1(self.bn1(self.up_conv(out)))
        xcat = torch.cat((out, skipx), 1)  # order (out, skipx): vnet.py:152
        out = self.ops(xcat)
        return self.relu2(out + xcat)


class OutputTransition(nn.Module):  # vnet.py:159-175
    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels, num_classes, kernel_size=5, padding=2)
        self.bn1 = PaddleBatchNorm3D(num_classes)
        self.conv2 = nn.Conv3d(num_classes, num_classes, kernel_size=1)
        self.relu1 = PaddlePReLU(num_classes)
        _paddle_conv_init(self.conv1)
        _paddle_conv_init(self.conv2)

    def forward(self, x):
        return self.conv2(self.relu1(self.bn1(self.conv1(x))))


class VNetOracle(nn.Module):
    """vnet.py:178-267.  forward(x, masks) -> [logits]; masks from make_dropout_masks (train mode only)."""

    def __init__(self, elu=False, in_channels=1, num_classes=4, pretrained=None,
                 kernel_size=((2, 2, 2),) * 4, stride_size=((2, 2, 2),) * 4):
        super().__init__()
        if elu:
            raise NotImplementedError("elu=True (nn.ELU) is not restated; the reference notes NaN grads (core/train.py:139)")
        self.best_loss = 1000000
        self.num_classes = num_classes
        self.in_channels = in_channels
        k, s = kernel_size, stride_size
        self.in_tr = InputTransition(in_channels)
        self.down_tr32 = DownTransition(16, 1, stride=s[0], kernel=k[0])
        self.down_tr64 = DownTransition(32, 2, stride=s[1], kernel=k[1])
        self.down_tr128 = DownTransition(64, 3, dropout=True, stride=s[2], kernel=k[2])
        self.down_tr256 = DownTransition(128, 2, dropout=True, stride=s[3], kernel=k[3])
        self.up_tr256 = UpTransition(256, 256, 2, dropout=True, dropout2=True, stride=s[3], kernel=k[3])
        self.up_tr128 = UpTransition(256, 128, 2, dropout=True, dropout2=True, stride=s[2], kernel=k[2])
        self.up_tr64 = UpTransition(128, 64, 1, stride=s[1], kernel=k[1])
        self.up_tr32 = UpTransition(64, 32, 1, stride=s[0], kernel=k[0])
        self.out_tr = OutputTransition(32, num_classes)
        self.pretrained = pretrained

    def forward(self, x, masks: Optional[Dict[str, torch.Tensor]] = None):  # vnet.py:256-267
        m = masks or {}
        if self.training and not masks:
            raise ValueError("train-mode forward needs explicit dropout masks (make_dropout_masks)")
        out16 = self.in_tr(x)
        out32 = self.down_tr32(out16)
        out64 = self.down_tr64(out32)
        out128 = self.down_tr128(out64, m.get("down_tr128"))
        out256 = self.down_tr256(out128, m.get("down_tr256"))
        out = self.up_tr256(out256, out128, m.get("up_tr256.x"), m.get("up_tr256.skip"))
        out = self.up_tr128(out, out64, m.get("up_tr128.x"), m.get("up_tr128.skip"))
        out = self.up_tr64(out, out32)
        out = self.up_tr32(out, out16)
        out = self.out_tr(out)
        return [out]


class VNetDeepSupOracle(nn.Module):
    """vnet_deepsup.py:176-281: the VNet trunk with three 3x3x3 deep-supervision heads on the decoder stages,
    trilinearly resized to the input size (F.interpolate(mode='trilinear'), Paddle defaults align_corners=False,
    align_mode=0 == torch's align_corners=False).  forward -> [out, d1 (256-ch stage), d2 (128), d3 (64)].
    `out_tr_all` is built but never used by the reference's forward (its parameters only exist in the state dict)."""

    def __init__(self, elu=False, in_channels=1, num_classes=4, pretrained=None,
                 kernel_size=((2, 2, 2),) * 4, stride_size=((2, 2, 2),) * 4):
        super().__init__()
        if elu:
            raise NotImplementedError("elu=True (nn.ELU) is not restated")
        self.best_loss = 1000000
        self.num_classes, self.in_channels, self.pretrained = num_classes, in_channels, pretrained
        k, s = kernel_size, stride_size
        self.in_tr = InputTransition(in_channels)
        self.down_tr32 = DownTransition(16, 1, stride=s[0], kernel=k[0])
        self.down_tr64 = DownTransition(32, 2, stride=s[1], kernel=k[1])
        self.down_tr128 = DownTransition(64, 3, dropout=True, stride=s[2], kernel=k[2])
        self.down_tr256 = DownTransition(128, 2, dropout=True, stride=s[3], kernel=k[3])
        self.up_tr256 = UpTransition(256, 256, 2, dropout=True, dropout2=True, stride=s[3], kernel=k[3])
        self.up_tr128 = UpTransition(256, 128, 2, dropout=True, dropout2=True, stride=s[2], kernel=k[2])
        self.up_tr64 = UpTransition(128, 64, 1, stride=s[1], kernel=k[1])
        self.up_tr32 = UpTransition(64, 32, 1, stride=s[0], kernel=k[0])
        self.out_tr32 = OutputTransition(32, num_classes)                      # vnet_deepsup.py:244
        self.out_tr64 = nn.Conv3d(64, num_classes, kernel_size=3, padding=1)   # :245
        self.out_tr128 = nn.Conv3d(128, num_classes, kernel_size=3, padding=1)
        self.out_tr256 = nn.Conv3d(256, num_classes, kernel_size=3, padding=1)
        for conv in (self.out_tr64, self.out_tr128, self.out_tr256):
            _paddle_conv_init(conv)
        self.out_tr_all = OutputTransition(4 * num_classes, num_classes)      # :248 (unused in forward)

    def forward(self, x, masks: Optional[Dict[str, torch.Tensor]] = None):  # vnet_deepsup.py:256-275
        m = masks or {}
        if self.training and not masks:
            raise ValueError("train-mode forward needs explicit dropout masks (make_dropout_masks)")
        size = x.shape[2:]
        out16 = self.in_tr(x)
        out32 = self.down_tr32(out16)
        out64 = self.down_tr64(out32)
        out128 = self.down_tr128(out64, m.get("down_tr128"))
        out256 = self.down_tr256(out128, m.get("down_tr256"))
        out = self.up_tr256(out256, out128, m.get("up_tr256.x"), m.get("up_tr256.skip"))
        d1 = F.interpolate(self.out_tr256(out), size=size, mode="trilinear", align_corners=False)
        out = self.up_tr128(out, out64, m.get("up_tr128.x"), m.get("up_tr128.skip"))
        d2 = F.interpolate(self.out_tr128(out), size=size, mode="trilinear", align_corners=False)
        out = self.up_tr64(out, out32)
        d3 = F.interpolate(self.out_tr64(out), size=size, mode="trilinear", align_corners=False)
        out = self.up_tr32(out, out16)
        out = self.out_tr32(out)
        return [out, d1, d2, d3]


def deepsup_losses():
    """vnetdeepsup_mri_spine_seg_512_512_12_15k.yml:12-20: one MixedLoss(CE, Dice) PER output (config.py:265-267
    repeats the single entry, _load_object builds four separate objects), coef 0.25 each"""
    return {"types": [MixedLoss([CrossEntropyLoss(), DiceLoss()], [1, 1]) for _ in range(4)],
            "coef": [0.25, 0.25, 0.25, 0.25]}


# --------------------------------------------------------------------------------------------
# Losses
# --------------------------------------------------------------------------------------------

def flatten(t):  # models/losses/loss_utils.py:18-28
    order = (1, 0) + tuple(range(2, t.dim()))
    return t.permute(*order).reshape(t.shape[1], -1)


def class_weights(logit):  # models/losses/loss_utils.py:31-40
    p = F.softmax(logit, dim=1)
    fl = flatten(p)
    w = (1.0 - fl).sum(-1) / fl.sum(-1)
    return w.detach()


class DiceLoss(nn.Module):  # models/losses/dice_loss.py:23-102
    def __init__(self, sigmoid_norm=True, weight=None):
        super().__init__()
        self.weight = weight
        self.sigmoid_norm = sigmoid_norm

    def forward(self, logits, labels):
        if logits.dim() == 4:
            logits = logits.unsqueeze(0)
        c = logits.shape[1]
        lab = labels.long()
        valid = ((lab >= 0) & (lab < c)).unsqueeze(1)
        one_hot = F.one_hot(lab.clamp(0, c - 1), c).permute(0, 4, 1, 2, 3).to(logits.dtype) * valid
        p = torch.sigmoid(logits) if self.sigmoid_norm else F.softmax(logits, dim=1)
        pi, ti = flatten(p), flatten(one_hot)
        intersect = (pi * ti).sum(-1)
        if self.weight is not None:
            intersect = self.weight * intersect
        denom = (pi * pi).sum(-1) + (ti * ti).sum(-1)
        dice = 2 * (intersect / denom.clamp(min=1e-6))  # dice_loss.py:62-74
        return 1.0 - dice.mean(), dice.detach().cpu().numpy()


class CrossEntropyLoss(nn.Module):  # models/losses/cross_entropy_loss.py:23-87
    def __init__(self, weight=None, ignore_index=255, data_format="NCDHW"):
        super().__init__()
        self.ignore_index = ignore_index
        self.EPS = 1e-8
        self.data_format = data_format
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32)

    def forward(self, logit, label):
        label = label.long()
        if logit.dim() == 4:
            logit = logit.unsqueeze(0)
        if self.weight is None:  # computed once from the FIRST logits, then cached (:68-69)
            self.weight = class_weights(logit)
        if logit.shape[1] != len(self.weight):
            raise ValueError("The number of weights = {} must be the same as the number of classes = {}.".format(
                len(self.weight), logit.shape[1]))
        return F.cross_entropy(logit + self.EPS, label, weight=self.weight.to(logit.dtype),
                               ignore_index=self.ignore_index, reduction="mean")


class MixedLoss(nn.Module):  # models/losses/mixes_losses.py:22-60
    def __init__(self, losses, coef):
        super().__init__()
        if not isinstance(losses, list):
            raise TypeError("`losses` must be a list!")
        if not isinstance(coef, list):
            raise TypeError("`coef` must be a list!")
        if len(losses) != len(coef):
            raise ValueError("The length of `losses` should equal to `coef`, but they are {} and {}.".format(
                len(losses), len(coef)))
        self.losses = losses
        self.coef = coef

    def forward(self, logits, labels):
        loss_list, per_channel_dice = [], None
        for i, loss in enumerate(self.losses):
            out = loss(logits, labels)
            if type(loss).__name__ == "DiceLoss":
                out, per_channel_dice = out
            loss_list.append(out * self.coef[i])
        return loss_list, per_channel_dice


def loss_computation(logits_list, labels, losses):  # utils/loss_utils.py:25-52
    if len(logits_list) != len(losses["types"]):
        raise RuntimeError("The length of logits_list should equal to the types of loss config: {} != {}.".format(
            len(logits_list), len(losses["types"])))
    loss_list, per_channel_dice = [], None
    for i, logits in enumerate(logits_list):
        loss_i, coef_i = losses["types"][i], losses["coef"][i]
        name = loss_i.__class__.__name__
        if name == "MixedLoss":
            mixed, per_channel_dice = loss_i(logits, labels)
            loss_list += [coef_i * m for m in mixed]
        elif name == "DiceLoss":
            loss, per_channel_dice = loss_i(logits, labels)
            loss_list.append(coef_i * loss)
        else:
            loss_list.append(coef_i * loss_i(logits, labels))
    return loss_list, per_channel_dice


def default_losses():
    """configs/lung_coronavirus/lung_coronavirus.yml:41-49 — MixedLoss([CE, Dice],[1,1]) x coef 1."""
    return {"types": [MixedLoss([CrossEntropyLoss(), DiceLoss()], [1, 1])], "coef": [1]}


# --------------------------------------------------------------------------------------------
# Optimizer / LR (cvlibs/config.py:156-169,203-232) and the train step (core/train.py:123-155)
# --------------------------------------------------------------------------------------------

class PolynomialDecay:
    def __init__(self, learning_rate, decay_steps, end_lr=0.0, power=0.9):
        self.base_lr, self.decay_steps, self.end_lr, self.power = learning_rate, decay_steps, end_lr, power
        self.last_epoch = 0

    def get_lr(self):
        t = min(self.last_epoch, self.decay_steps)
        return (self.base_lr - self.end_lr) * (1 - t / self.decay_steps) ** self.power + self.end_lr

    def step(self):
        self.last_epoch += 1


class Momentum:
    def __init__(self, lr, params: Sequence[torch.Tensor], momentum=0.9, weight_decay=0.0):
        self.lr = lr
        self.params = list(params)
        self.mu = momentum
        self.wd = float(weight_decay or 0.0)
        self.velocity = [torch.zeros_like(p) for p in self.params]

    def get_lr(self):
        return self.lr.get_lr() if isinstance(self.lr, PolynomialDecay) else float(self.lr)

    @torch.no_grad()
    def step(self):
        lr = self.get_lr()
        for p, v in zip(self.params, self.velocity):
            if p.grad is None:
                continue
            g = p.grad + self.wd * p
            v.mul_(self.mu).add_(g)
            p.sub_(lr * v)

    def clear_grad(self):
        for p in self.params:
            p.grad = None


def train_step(model: VNetOracle, losses, opt: Momentum, images, labels, masks):
    """One iteration of core/train.py:123-155; returns (loss float, loss_list floats, per_channel_dice)."""
    logits_list = model(images, masks)
    loss_list, dice = loss_computation(logits_list, labels.to(torch.int32), losses)
    loss = sum(loss_list)
    loss.backward()
    opt.step()
    if isinstance(opt.lr, PolynomialDecay):
        opt.lr.step()
    opt.clear_grad()
    return float(loss.detach()), [float(l.detach()) for l in loss_list], dice


# --------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench (SURVEY.md §8d cfg 2)
# --------------------------------------------------------------------------------------------

def synthetic_batch(batch: int, shape: Tuple[int, int, int], num_classes: int, seed: int = 0):
    """images f32 [N,1,D,H,W] in [0,1] (per-volume /max as transforms/transform.py:67-69);
    labels int32 [N,D,H,W] from smoothed noise so classes are spatially coherent."""
    g = torch.Generator().manual_seed(seed)
    d, h, w = shape
    img = torch.rand(batch, 1, d, h, w, generator=g)
    img = img / img.amax(dim=(1, 2, 3, 4), keepdim=True)
    g2 = torch.Generator().manual_seed(seed + 1)
    noise = torch.rand(batch, 1, d, h, w, generator=g2)
    k = 5
    smooth = noise
    for _ in range(2):
        smooth = F.avg_pool3d(F.pad(smooth, (k // 2,) * 6, mode="replicate"), k, stride=1)
    flat = smooth.flatten(1)
    qs = torch.quantile(flat[:, :: max(1, flat.shape[1] // 65536)], torch.linspace(0, 1, num_classes + 1)[1:-1], dim=1)
    lab = torch.zeros(batch, d, h, w, dtype=torch.int32)
    for c in range(num_classes - 1):
        lab += (smooth[:, 0] > qs[c].view(-1, 1, 1, 1)).to(torch.int32)
    return img, lab
