"""medicalseg/core/__init__.py:15-17 of the reference."""
from .train import train  # noqa: F401
from .val import evaluate  # noqa: F401
from . import infer  # noqa: F401
