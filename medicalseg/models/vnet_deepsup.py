"""medicalseg.models.vnet_deepsup.VNetDeepSup (reference vnet_deepsup.py:176-281) -> the sm_100a engine."""
from medicalseg_b200.models.vnet_deepsup import VNetDeepSup  # noqa: F401
