"""Datasets (reference: medicalseg/datasets/dataset.py:28-125).  The .npy reader keeps the reference contract —
`<dataset_root>/{train,val}_list.txt` with "image.npy label.npy" pairs, items are (im [1,D,H,W] f32 scaled by its
max as transforms/transform.py:67-69, label [D,H,W] int, path).  A `transforms` list (configs: RandomResizedCrop3D,
RandomRotation3D, RandomFlip3D ...) is wrapped in `transforms.Compose` as dataset.py:113 does and runs ON THE DEVICE:
the item then comes back as CUDA tensors.  `SyntheticVolumes` feeds the benchmark / smoke configurations."""
from __future__ import annotations

import os

import numpy as np
import torch


class SyntheticVolumes:
    def __init__(self, num_classes=2, shape=(128, 128, 128), length=16, seed=0, mode="train", **_):
        self.num_classes, self.shape, self.length, self.seed, self.mode = num_classes, tuple(shape), length, seed, mode
        self.ignore_index = 255
        self.transforms = None

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        d, h, w = self.shape
        img = torch.rand(1, d, h, w, generator=g)
        img = img / img.max()
        low = torch.rand(1, 1, max(d // 8, 1), max(h // 8, 1), max(w // 8, 1), generator=g)
        sm = torch.nn.functional.interpolate(low, size=self.shape, mode="trilinear", align_corners=False)[0, 0]
        qs = torch.quantile(sm.flatten()[::max(1, sm.numel() // 65536)],
                            torch.linspace(0, 1, self.num_classes + 1)[1:-1])
        lab = torch.zeros(d, h, w, dtype=torch.int32)
        for q in qs:
            lab += (sm > q).to(torch.int32)
        return img, lab, "synthetic_%d" % idx



class NpyVolumeDataset:
    def __init__(self, dataset_root, result_dir=None, transforms=None, num_classes=None, mode="train",
                 ignore_index=255, dataset_json_path="", **_):
        self.dataset_root, self.result_dir, self.mode = dataset_root, result_dir, mode.lower()
        self.num_classes, self.ignore_index, self.dataset_json_path = num_classes, ignore_index, dataset_json_path
        self.transforms = None
        if transforms:  # dataset.py:113: T.Compose(transforms); an empty list means "only read the volumes"
            from .transforms import Compose
            self.transforms = transforms if isinstance(transforms, Compose) else Compose(list(transforms))
        if self.mode not in ("train", "val"):
            raise ValueError("`mode` should be 'train' or 'val', but got {}.".format(mode))
        if num_classes is None:
            raise ValueError("`num_classes` is necessary, but it is None.")
        if not os.path.exists(dataset_root):
            raise FileNotFoundError("there is not `dataset_root`: {}.".format(dataset_root))
        self.file_list = []
        with open(os.path.join(dataset_root, "%s_list.txt" % self.mode)) as f:
            for line in f:
                items = line.strip().split()
                if len(items) != 2:
                    raise Exception("File list format incorrect! It should be image_name label_name\\n")
                self.file_list.append([os.path.join(dataset_root, items[0]), os.path.join(dataset_root, items[1])])
        if self.mode == "train":
            self.file_list = self.file_list * 10  # dataset.py:110-111

    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        image_path, label_path = self.file_list[idx]
        if self.transforms is not None:
            im, label = self.transforms(image_path, label_path)  # device tensors: [1,D,H,W] f32 / max, [D,H,W] i32
            return im, label, image_path
        im = np.load(image_path).astype(np.float32)
        label = np.load(label_path)
        im = im[None]
        if im.max() > 0:
            im = im / im.max()
        return torch.from_numpy(im), torch.from_numpy(label.astype(np.int32)), image_path
