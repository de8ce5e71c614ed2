"""medicalseg.transforms.transform (reference transform.py:27-339) -> device-side transforms."""
from medicalseg_b200.transforms import Compose, RandomFlip3D, RandomResizedCrop3D, RandomRotation3D, Resize3D  # noqa: F401
