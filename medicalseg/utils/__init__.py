"""medicalseg/utils/__init__.py:15-25 of the reference - the names train.py / val.py import (`get_sys_env`, `logger`,
`config_check`, `utils`, `loss_computation`).  Download / VisualDL / FLOP-counting
helpers are outside the hot path (DESIGN.md 'Out of scope')."""
from . import logger  # noqa: F401
from . import utils  # noqa: F401
from .utils import load_entire_model, load_pretrained_model, resume  # noqa: F401
from .config_check import config_check  # noqa: F401
from .env_util import get_sys_env  # noqa: F401
from medicalseg_b200.models.losses import check_logits_losses, loss_computation  # noqa: F401
