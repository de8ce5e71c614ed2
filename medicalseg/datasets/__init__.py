"""medicalseg/datasets/__init__.py:15-17 of the reference: the .npy list reader under the reference class names."""
from medicalseg_b200.datasets import NpyVolumeDataset as MedicalDataset  # noqa: F401
from medicalseg_b200.datasets import NpyVolumeDataset as LungCoronavirus  # noqa: F401
from medicalseg_b200.datasets import NpyVolumeDataset as MRISpineSeg  # noqa: F401
