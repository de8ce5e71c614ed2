"""medicalseg.core.infer (reference core/infer.py:20-94)."""
from medicalseg_b200.core import get_reverse_list, inference, reverse_transform  # noqa: F401
