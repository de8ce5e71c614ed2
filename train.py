#!/usr/bin/env python
"""Entry point with the reference's flags (train.py:26-115,118-189 of PaddleCV-SIG/MedicalSeg) on the B200 path.
Multi-GPU: `torchrun --nproc-per-node N --master-addr 127.0.0.1 train.py --config ...` (one process per GPU)."""
import argparse
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse_args():
    p = argparse.ArgumentParser(description="Model training")
    p.add_argument("--config", dest="cfg", help="The config file.", default=None, type=str)
    p.add_argument("--iters", dest="iters", help="iters for training", type=int, default=None)
    p.add_argument("--batch_size", dest="batch_size", help="Mini batch size of one gpu or cpu", type=int, default=None)
    p.add_argument("--learning_rate", dest="learning_rate", help="Learning rate", type=float, default=None)
    p.add_argument("--save_interval", dest="save_interval", type=int, default=1000)
    p.add_argument("--resume_model", dest="resume_model", type=str, default=None)
    p.add_argument("--save_dir", dest="save_dir", type=str, default="./output")
    p.add_argument("--keep_checkpoint_max", dest="keep_checkpoint_max", type=int, default=5)
    p.add_argument("--num_workers", dest="num_workers", type=int, default=0)
    p.add_argument("--do_eval", dest="do_eval", action="store_true")
    p.add_argument("--log_iters", dest="log_iters", type=int, default=100)
    p.add_argument("--use_vdl", dest="use_vdl", action="store_true")
    p.add_argument("--seed", dest="seed", type=int, default=None)
    p.add_argument("--data_format", dest="data_format", type=str, default="NCHW")
    p.add_argument("--profiler_options", type=str, default=None)
    p.add_argument("--to_static_training", action="store_true")
    return p.parse_args()


def main(args):
    if args.seed is not None:
        torch.manual_seed(args.seed); np.random.seed(args.seed); random.seed(args.seed)
    if not torch.cuda.is_available():
        raise RuntimeError("train.py needs a CUDA device: the B200 path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not args.cfg:
        raise RuntimeError("No configuration file specified.")
    from medicalseg_b200.cvlibs import Config
    from medicalseg_b200.core import train
    cfg = Config(args.cfg, learning_rate=args.learning_rate, iters=args.iters, batch_size=args.batch_size)
    train_dataset = cfg.train_dataset
    if train_dataset is None:
        raise RuntimeError("The training dataset is not specified in the configuration file.")
    if len(train_dataset) == 0:
        raise ValueError("The length of train_dataset is 0. Please check if your dataset is valid")
    val_dataset = cfg.val_dataset if args.do_eval else None
    losses = cfg.loss
    print("------------Environment Information-------------\n" + str(cfg) + "------------------------------------------------")
    train(cfg.model, train_dataset, val_dataset=val_dataset, optimizer=cfg.optimizer, save_dir=args.save_dir,
          iters=cfg.iters, batch_size=cfg.batch_size, resume_model=args.resume_model,
          save_interval=args.save_interval, log_iters=args.log_iters, num_workers=args.num_workers,
          use_vdl=args.use_vdl, losses=losses, keep_checkpoint_max=args.keep_checkpoint_max,
          profiler_options=args.profiler_options, to_static_training=args.to_static_training,
          seed=args.seed or 0)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(parse_args())
