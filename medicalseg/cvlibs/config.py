"""medicalseg.cvlibs.config.Config (reference config.py:29-429)."""
from medicalseg_b200.cvlibs import Config  # noqa: F401
