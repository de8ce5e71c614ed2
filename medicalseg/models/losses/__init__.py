"""medicalseg/models/losses/__init__.py:14-18 of the reference.  BCELoss is outside the VNet hot path and is
not provided; asking for it raises with that explanation instead of an AttributeError."""
from medicalseg_b200.models.losses import (CrossEntropyLoss, DiceLoss, MixedLoss, class_weights,  # noqa: F401
                                           loss_computation)

__all__ = ["CrossEntropyLoss", "DiceLoss", "MixedLoss", "class_weights", "flatten"]


def flatten(tensor):
    """models/losses/loss_utils.py:18-28: (N, C, D, H, W) -> (C, N*D*H*W) (a view permutation; no kernel)"""
    c = tensor.shape[1]
    return tensor.movedim(1, 0).reshape(c, -1)


def __getattr__(name):
    if name == "BCELoss":
        raise NotImplementedError("BCELoss is not on the VNet hot path this package accelerates (configs use "
                                  "MixedLoss[CrossEntropyLoss, DiceLoss]); see DESIGN.md 'Out of scope'")
    raise AttributeError(name)
