"""Momentum + L2 and PolynomialDecay with Paddle semantics (reference: cvlibs/config.py:156-169,203-232;
step order core/train.py:140-151), as ONE fused multi-tensor kernel over the model's flat parameter buffer."""
from __future__ import annotations

import torch

from . import ops


class PolynomialDecay:
    """paddle.optimizer.lr.PolynomialDecay(cycle=False): lr = (lr0-end)*(1-min(t,T)/T)**power + end"""

    def __init__(self, learning_rate, decay_steps, end_lr=0.0, power=0.9, **_):
        self.base_lr, self.decay_steps, self.end_lr, self.power = learning_rate, decay_steps, end_lr, power
        self.last_epoch = 0

    def get_lr(self):
        t = min(self.last_epoch, self.decay_steps)
        return (self.base_lr - self.end_lr) * (1 - t / self.decay_steps) ** self.power + self.end_lr

    def __call__(self):
        return self.get_lr()

    def step(self):
        self.last_epoch += 1

    def state_dict(self):
        return {"last_epoch": self.last_epoch}

    def set_state_dict(self, sd):
        self.last_epoch = int(sd["last_epoch"])


class Momentum:
    """paddle.optimizer.Momentum(learning_rate, parameters, momentum, weight_decay) — g += wd*p; v = mu*v + g;
    p -= lr*v, applied to EVERY parameter (BN scale/shift, PReLU slopes and biases included, as in the reference)."""

    def __init__(self, learning_rate, parameters, momentum=0.9, weight_decay=0.0, grad_scale=1.0, **_):
        store = getattr(parameters, "store", None)
        if store is None:
            raise TypeError("Momentum needs model.parameters() of a medicalseg_b200 model (flat parameter store)")
        self._learning_rate = learning_rate
        self._store, self._owner = store, getattr(parameters, "owner", None)
        self.momentum, self.weight_decay, self.grad_scale = momentum, float(weight_decay or 0.0), grad_scale
        self.velocity = torch.zeros_like(store.flat)
        self.lr_dev = None  # device copy of the learning rate (set by GraphedTrainStep: graph replays read it)

    def get_lr(self):
        lr = self._learning_rate
        return lr.get_lr() if hasattr(lr, "get_lr") else float(lr)

    def step(self):
        if self.lr_dev is not None:
            ops.momentum_step_lrdev(self._store.flat, self._store.grad, self.velocity, self.lr_dev, self.momentum,
                                    self.weight_decay, self.grad_scale)
        else:
            ops.momentum_step(self._store.flat, self._store.grad, self.velocity, self.get_lr(), self.momentum,
                              self.weight_decay, self.grad_scale)
        if self._owner is not None:
            self._owner.mark_parameters_updated()

    def clear_grad(self):
        ops.zero_(self._store.grad)

    def state_dict(self):
        # per-parameter momentum in the reference (Paddle) shapes: independent of the internal (tap-major) layout
        sd = {"velocity": {name: self._store.view_of(self.velocity, name).detach().contiguous().clone()
                           for name, slot in self._store.slots.items() if not slot.is_buffer}}
        if hasattr(self._learning_rate, "state_dict"):
            sd["LR_Scheduler"] = self._learning_rate.state_dict()
        return sd

    def set_state_dict(self, sd):
        vel = sd["velocity"]
        if torch.is_tensor(vel):  # legacy flat buffer (same layout only)
            self.velocity.copy_(vel)
        else:
            for name, v in vel.items():
                self._store.view_of(self.velocity, name).copy_(torch.as_tensor(v).to(self.velocity.device))
        if "LR_Scheduler" in sd and hasattr(self._learning_rate, "set_state_dict"):
            self._learning_rate.set_state_dict(sd["LR_Scheduler"])
