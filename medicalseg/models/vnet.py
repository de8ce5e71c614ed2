"""medicalseg.models.vnet.VNet (reference vnet.py:178-267) -> the sm_100a engine."""
from medicalseg_b200.models.vnet import VNet  # noqa: F401
