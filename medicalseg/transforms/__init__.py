"""medicalseg/transforms/__init__.py:15-16 of the reference."""
from .transform import Compose, RandomFlip3D, RandomResizedCrop3D, RandomRotation3D, Resize3D  # noqa: F401
from . import functional  # noqa: F401
