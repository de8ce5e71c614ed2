"""medicalseg.utils.config_check (reference config_check.py:18-60): `num_classes` has to agree between the model
section of the YAML and the datasets; the agreed value is written back to both datasets."""


def _declared_num_classes(cfg, datasets):
    values = {ds.num_classes for ds in datasets if ds and hasattr(ds, "num_classes")}
    model_section = cfg.dic.get("model") or {}
    if model_section.get("num_classes"):
        values.add(model_section["num_classes"])
    return values


def num_classes_check(cfg, train_dataset, val_dataset):
    datasets = (train_dataset, val_dataset)
    values = _declared_num_classes(cfg, datasets)
    if not cfg.train_dataset and not cfg.val_dataset:
        raise ValueError("One of `train_dataset` or `val_dataset should be given, but there are none.")
    if len(values) != 1:
        if not values:
            raise ValueError("`num_classes` is not found. Please set it in model, train_dataset or val_dataset")
        raise ValueError("`num_classes` is not consistent: {}. Please set it consistently in model or train_dataset or "
                         "val_dataset".format(values))
    (agreed,) = values
    for ds in datasets:
        if ds:
            ds.num_classes = agreed


def config_check(cfg, train_dataset=None, val_dataset=None):
    num_classes_check(cfg, train_dataset, val_dataset)
