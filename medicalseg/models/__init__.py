"""medicalseg/models/__init__.py:15-17 of the reference: losses, VNet, VNetDeepSup."""
from .losses import *  # noqa: F401,F403
from .vnet import VNet  # noqa: F401
from .vnet_deepsup import VNetDeepSup  # noqa: F401
