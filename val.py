#!/usr/bin/env python
"""Evaluation entry point with the reference's flags (val.py:25-121)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse_args():
    p = argparse.ArgumentParser(description="Model evaluation")
    p.add_argument("--config", dest="cfg", help="The config file.", default=None, type=str)
    p.add_argument("--model_path", dest="model_path", help="The path of model for evaluation", type=str, default=None)
    p.add_argument("--save_dir", dest="save_dir", type=str, default="saved_model/vnet_lung_coronavirus_128_128_128_15k")
    p.add_argument("--num_workers", dest="num_workers", type=int, default=0)
    p.add_argument("--print_detail", dest="print_detail", action="store_true", default=True)
    p.add_argument("--use_vdl", dest="use_vdl", action="store_true")
    p.add_argument("--auc_roc", dest="add auc_roc metric", action="store_true")
    return p.parse_args()


def main(args):
    if not torch.cuda.is_available():
        raise RuntimeError("val.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if not args.cfg:
        raise RuntimeError("No configuration file specified.")
    from medicalseg_b200.cvlibs import Config
    from medicalseg_b200.core import evaluate
    from medicalseg_b200.utils import load_entire_model
    cfg = Config(args.cfg)
    val_dataset = cfg.val_dataset
    if val_dataset is None:
        raise RuntimeError("The verification dataset is not specified in the configuration file.")
    if len(val_dataset) == 0:
        raise ValueError("The length of val_dataset is 0. Please check if your dataset is valid")
    model = cfg.model
    if args.model_path:
        load_entire_model(model, args.model_path)
        print("Loaded trained params of model successfully")
    evaluate(model, val_dataset, cfg.loss, num_workers=args.num_workers, print_detail=args.print_detail,
             save_dir=args.save_dir)


if __name__ == "__main__":
    main(parse_args())
