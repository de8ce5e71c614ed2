"""Component registry + YAML Config with `_base_` inheritance (reference: medicalseg/cvlibs/manager.py:23-149,
config.py:29-429).  Same semantics: components are looked up by class name from `type:`, `_inherited_: False`
opts a sub-dict out of the merge, CLI learning_rate / batch_size / iters override the file, loss `types` are
broadcast over `coef`, `num_classes` is injected from the dataset when the model section omits it.
With more than one rank the model is built with `sync_bn=True` - the reference converts every BatchNorm to
SyncBatchNorm there (config.py:322) - unless the YAML says `sync_bn: false` (per-rank statistics)."""
from __future__ import annotations

import codecs
import inspect
import os
from collections.abc import Sequence
from typing import Any

import yaml


class ComponentManager:
    def __init__(self, name=None):
        self._components_dict = dict()
        self._name = name

    def __len__(self):
        return len(self._components_dict)

    def __repr__(self):
        return "{}:{}".format(self._name or self.__class__.__name__, list(self._components_dict.keys()))

    def __getitem__(self, item):
        if item not in self._components_dict:
            raise KeyError("{} does not exist in availabel {}".format(item, self))
        return self._components_dict[item]

    @property
    def components_dict(self):
        return self._components_dict

    def _add_single_component(self, component):
        if not (inspect.isclass(component) or inspect.isfunction(component)):
            raise TypeError("Expect class/function type, but received {}".format(type(component)))
        name = component.__name__
        if name in self._components_dict:
            raise KeyError("{} exists already!".format(name))
        self._components_dict[name] = component

    def add_component(self, components):
        if isinstance(components, Sequence):
            for c in components:
                self._add_single_component(c)
        else:
            self._add_single_component(components)
        return components


MODELS = ComponentManager("models")
BACKBONES = ComponentManager("backbones")
DATASETS = ComponentManager("datasets")
TRANSFORMS = ComponentManager("transforms")
LOSSES = ComponentManager("losses")


def _register_defaults():
    from .models import VNet, VNetDeepSup
    from .models.losses import CrossEntropyLoss, DiceLoss, MixedLoss
    from .datasets import NpyVolumeDataset, SyntheticVolumes
    from .transforms import Compose, RandomFlip3D, RandomResizedCrop3D, RandomRotation3D, Resize3D
    for mgr, comps in ((MODELS, [VNet, VNetDeepSup]), (LOSSES, [CrossEntropyLoss, DiceLoss, MixedLoss]),
                       (DATASETS, [NpyVolumeDataset, SyntheticVolumes]),
                       (TRANSFORMS, [Compose, RandomFlip3D, RandomResizedCrop3D, RandomRotation3D, Resize3D])):
        for c in comps:
            if c.__name__ not in mgr.components_dict:
                mgr.add_component(c)
    # reference dataset class names resolve to the .npy list reader
    for alias in ("LungCoronavirus", "MRISpineSeg"):
        if alias not in DATASETS.components_dict:
            DATASETS.components_dict[alias] = NpyVolumeDataset


class Config:
    def __init__(self, path: str, learning_rate: float = None, batch_size: int = None, iters: int = None):
        if not path:
            raise ValueError("Please specify the configuration file path.")
        if not os.path.exists(path):
            raise FileNotFoundError("File {} does not exist".format(path))
        if not (path.endswith("yml") or path.endswith("yaml")):
            raise RuntimeError("Config file should in yaml format!")
        _register_defaults()
        self.dic = self._parse_from_yaml(path)
        self._model = None
        self._losses = None
        self._train_dataset = self._val_dataset = None
        self.update(learning_rate=learning_rate, batch_size=batch_size, iters=iters)

    def _update_dic(self, dic, base_dic):
        base_dic, dic = base_dic.copy(), dic.copy()
        if dic.get("_inherited_", True) is False:
            dic.pop("_inherited_")
            return dic
        for key, val in dic.items():
            if isinstance(val, dict) and key in base_dic:
                base_dic[key] = self._update_dic(val, base_dic[key])
            else:
                base_dic[key] = val
        return base_dic

    def _parse_from_yaml(self, path: str):
        with codecs.open(path, "r", "utf-8") as file:
            dic = yaml.load(file, Loader=yaml.FullLoader)
        if "_base_" in dic:
            base_path = os.path.join(os.path.dirname(path), dic.pop("_base_"))
            dic = self._update_dic(dic, self._parse_from_yaml(base_path))
        return dic

    def update(self, learning_rate=None, batch_size=None, iters=None):
        if learning_rate:
            if "lr_scheduler" in self.dic:
                self.dic["lr_scheduler"]["learning_rate"] = learning_rate
            else:
                self.dic.setdefault("learning_rate", {})["value"] = learning_rate
        if batch_size:
            self.dic["batch_size"] = batch_size
        if iters:
            self.dic["iters"] = iters

    @property
    def batch_size(self) -> int:
        return self.dic.get("batch_size", 1)

    @property
    def iters(self) -> int:
        iters = self.dic.get("iters")
        if not iters:
            raise RuntimeError("No iters specified in the configuration file.")
        return iters

    @property
    def lr_scheduler(self):
        from .optimizer import PolynomialDecay
        if "lr_scheduler" not in self.dic:
            raise RuntimeError("No `lr_scheduler` specified in the configuration file.")
        params = self.dic.get("lr_scheduler").copy()
        lr_type = params.pop("type")
        if lr_type != "PolynomialDecay":
            raise RuntimeError("Only PolynomialDecay is implemented (reference configs use it); got %s" % lr_type)
        params.setdefault("decay_steps", self.iters)
        params.setdefault("end_lr", 0)
        params.setdefault("power", 0.9)
        return PolynomialDecay(**params)

    @property
    def optimizer(self):
        from .optimizer import Momentum
        args = self.dic.get("optimizer", {}).copy()
        opt_type = args.pop("type", "sgd")
        if opt_type != "sgd":
            raise RuntimeError("Only the `sgd` (Momentum) optimizer is implemented; got {}".format(opt_type))
        args.setdefault("momentum", 0.9)
        return Momentum(self.lr_scheduler, parameters=self.model.parameters(), **args)

    @property
    def loss(self) -> dict:
        if self._losses is None:
            args = self.dic.get("loss", {}).copy()
            if not ("types" in args and "coef" in args):
                raise ValueError('Loss config should contain keys of "types" and "coef"')
            if len(args["types"]) != len(args["coef"]):
                if len(args["types"]) == 1:
                    args["types"] = args["types"] * len(args["coef"])
                else:
                    raise ValueError("The length of types should equal to coef or equal to 1 in loss config, but they "
                                     "are {} and {}.".format(len(args["types"]), len(args["coef"])))
            self._losses = {"types": [self._load_object(item) for item in args["types"]], "coef": args["coef"]}
        return self._losses

    @property
    def model(self):
        model_cfg = (self.dic.get("model") or {}).copy()
        if not model_cfg:
            raise RuntimeError("No model specified in the configuration file.")
        if "num_classes" not in model_cfg:
            ds = self.train_dataset if self.dic.get("train_dataset") else self.val_dataset
            if ds is not None and hasattr(ds, "num_classes"):
                model_cfg["num_classes"] = ds.num_classes
        if self._model is None:
            component = self._load_component(model_cfg["type"])
            import torch.distributed as dist
            if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                    and "sync_bn" not in model_cfg and "sync_bn" in inspect.signature(component).parameters):
                model_cfg["sync_bn"] = True  # config.py:322: paddle.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            self._model = self._load_object(model_cfg)
        return self._model

    @property
    def train_dataset(self):
        if self._train_dataset is None and self.dic.get("train_dataset"):
            self._train_dataset = self._load_object(self.dic["train_dataset"].copy())
        return self._train_dataset

    @property
    def val_dataset(self):
        if self._val_dataset is None and self.dic.get("val_dataset"):
            self._val_dataset = self._load_object(self.dic["val_dataset"].copy())
        return self._val_dataset

    def _load_component(self, com_name: str) -> Any:
        for com in (MODELS, BACKBONES, DATASETS, TRANSFORMS, LOSSES):
            if com_name in com.components_dict:
                return com[com_name]
        raise RuntimeError("The specified component was not found {}.".format(com_name))

    def _is_meta_type(self, item: Any) -> bool:
        return isinstance(item, dict) and "type" in item

    def _load_object(self, cfg: dict) -> Any:
        cfg = cfg.copy()
        if "type" not in cfg:
            raise RuntimeError("No object information in {}.".format(cfg))
        component = self._load_component(cfg.pop("type"))
        params = {}
        for key, val in cfg.items():
            if self._is_meta_type(val):
                params[key] = self._load_object(val)
            elif isinstance(val, list):
                params[key] = [self._load_object(i) if self._is_meta_type(i) else i for i in val]
            else:
                params[key] = val
        return component(**params)

    def __str__(self) -> str:
        return yaml.dump(self.dic)
