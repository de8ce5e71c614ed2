"""medicalseg.transforms.functional (reference functional.py:25-110) -> device kernels."""
from medicalseg_b200.transforms import flip_3d, resize_3d, resized_crop_3d, rotate_3d  # noqa: F401
