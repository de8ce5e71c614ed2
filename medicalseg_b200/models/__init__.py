from .vnet import VNet  # noqa: F401
from .vnet_deepsup import VNetDeepSup  # noqa: F401
from .losses import CrossEntropyLoss, DiceLoss, MixedLoss  # noqa: F401
