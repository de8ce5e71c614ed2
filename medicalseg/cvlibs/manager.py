"""medicalseg.cvlibs.manager (reference manager.py:23-149): the component registries, pre-populated with the B200
implementations so that `@manager.MODELS.add_component` user classes and YAML `type:` look-ups share one table."""
from medicalseg_b200.cvlibs import (BACKBONES, DATASETS, LOSSES, MODELS, TRANSFORMS, ComponentManager,  # noqa: F401
                                    _register_defaults)

_register_defaults()
