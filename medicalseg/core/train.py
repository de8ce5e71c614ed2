"""medicalseg.core.train.train (reference core/train.py:30-274)."""
from medicalseg_b200.core import train  # noqa: F401
