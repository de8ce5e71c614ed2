"""medicalseg/cvlibs/__init__.py:15-16 of the reference."""
from . import manager  # noqa: F401
from .config import Config  # noqa: F401
