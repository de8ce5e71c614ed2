"""medicalseg.core.val.evaluate (reference core/val.py:29-187)."""
from medicalseg_b200.core import evaluate  # noqa: F401
