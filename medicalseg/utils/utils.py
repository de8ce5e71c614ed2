"""medicalseg.utils.utils (reference utils.py:40-135): checkpoint helpers."""
from medicalseg_b200.utils import export_pdparams, load_entire_model, resume, save_checkpoint  # noqa: F401

load_pretrained_model = load_entire_model  # utils.py:76-112: same matching-shape load with per-key warnings
