"""Import shim: `medicalseg` resolves to the B200-native implementation (medicalseg_b200).

Reference-side code - `from medicalseg.models import VNet`, `from medicalseg.cvlibs import manager, Config`,
`from medicalseg.core import train, evaluate`, YAML `type: VNet` through the component registry - runs unchanged on
the sm_100a path when this repository precedes the reference checkout on `sys.path` (reference package layout:
medicalseg/__init__.py:15, models/__init__.py:15-17, cvlibs/__init__.py:15-16, core/__init__.py:15-17).
Every name re-exported here is the medicalseg_b200 object; nothing is computed in this package.
"""
from . import models, datasets, transforms, utils  # noqa: F401  (medicalseg/__init__.py:15)
from . import cvlibs, core  # noqa: F401

__version__ = "2.0.0-b200"
