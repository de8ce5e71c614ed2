"""Train / evaluate loops (reference: medicalseg/core/train.py:30-274, val.py:29-187, infer.py:62-94) on the B200
path: same iteration order (forward -> loss_computation -> backward -> optimizer.step -> lr step -> clear
gradients), same log line, same checkpoint layout (iter_N/model.pdparams + model.pdopt, best_model/).  Multi-GPU =
one process per GPU (torchrun), volumes sharded per rank, one bucketed NCCL gradient all-reduce per step."""
from __future__ import annotations

import os
import shutil
import time
from collections import deque

import numpy as np
import torch
import torch.distributed as dist

from .models.losses import loss_computation
from .parallel import DistributedGradReducer
from .utils import resume, save_checkpoint


def _log(msg, level="INFO"):
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        print("{} [{}]\t{}".format(time.strftime("%Y-%m-%d %H:%M:%S"), level, msg), flush=True)


def _collate(dataset, indices, device):
    ims, labs = zip(*[dataset[i][:2] for i in indices])
    return torch.stack(ims).to(device, non_blocking=True), torch.stack(labs).to(device, non_blocking=True)


def get_reverse_list(ori_shape, transforms):
    """core/infer.py:20-41: the shapes to restore, one ('resize', (d, h, w)) entry per Resize3D in `transforms`"""
    reverse_list = []
    d, h, w = ori_shape[0], ori_shape[1], ori_shape[2]
    for op in transforms or []:
        if op.__class__.__name__ in ["Resize3D"]:
            reverse_list.append(("resize", (d, h, w)))
            d, h, w = op.size[0], op.size[1], op.size[2]
    return reverse_list


def reverse_transform(logit, ori_shape, transforms):
    """core/infer.py:44-60: resize the logits back through every Resize3D of the validation transforms.  The reference
    passes mode='bilinear' for these 5-D tensors (infer.py:90), which Paddle rejects; the evident
    intent - linear interpolation of the logits, F.interpolate defaults = half-pixel centres - is what the trilinear
    kernel (msb_trilinear_fwd) computes."""
    from . import ops
    for kind, shape in get_reverse_list(ori_shape, transforms)[::-1]:
        if kind != "resize":
            raise Exception("Unexpected info '{}' in im_info".format(kind))
        out = torch.empty((*logit.shape[:2], *[int(v) for v in shape]), dtype=torch.float32, device=logit.device)
        ops.trilinear_fwd(logit.contiguous().float(), out)
        logit = out
    return logit


def inference(model, im, ori_shape=None, transforms=None):
    """core/infer.py:62-94 -> (pred int32 [N,1,D,H,W], logit [N,C,D,H,W])"""
    from . import ops
    logits = model(im)
    if not isinstance(logits, (list, tuple)):
        raise TypeError("The type of logits must be one of collections.abc.Sequence, e.g. list, tuple. But received {}"
                        .format(type(logits)))
    logit = logits[0]
    if ori_shape is not None and tuple(int(v) for v in ori_shape) != tuple(logit.shape[2:]):
        logit = reverse_transform(logit, ori_shape, transforms)
        if tuple(int(v) for v in ori_shape) != tuple(logit.shape[2:]):
            raise ValueError("inference(): the logits have shape {} but ori_shape is {} and the transforms hold no "
                             "Resize3D that explains the difference".format(tuple(logit.shape[2:]), tuple(ori_shape)))
    pred = torch.empty((logit.shape[0], 1, *logit.shape[2:]), dtype=torch.int32, device=logit.device)
    ops.argmax_channels(logit.contiguous().float(), pred)
    return pred, logit


def reduce_eval_metrics(mdice_sum, loss_sum, channel_dice_sum, n_local, device, num_classes=1):
    """per-rank sums over this rank's shard of the validation set -> GLOBAL means (mDice, loss, per-class Dice).  The
    reference does not reduce its metrics across ranks (core/val.py:58-66,168: every rank reports its own shard); here
    the sums and the sample count are all-reduced so every rank reports - and selects `best_model` on - the same figure."""
    if channel_dice_sum is None:  # a rank whose shard is empty (fewer validation volumes than ranks)
        channel_dice_sum = np.zeros(num_classes)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([mdice_sum, loss_sum, float(n_local)], device=device, dtype=torch.float64)
        cd = torch.as_tensor(np.asarray(channel_dice_sum), device=device, dtype=torch.float64)
        dist.all_reduce(t)
        dist.all_reduce(cd)
        mdice_sum, loss_sum, n_local = float(t[0]), float(t[1]), float(t[2])
        channel_dice_sum = cd.cpu().numpy()
    n = max(float(n_local), 1.0)
    return mdice_sum / n, loss_sum / n, np.asarray(channel_dice_sum) / n


def evaluate(model, eval_dataset, losses, num_workers=0, print_detail=True, save_dir=None, **_):
    new_loss = {"types": [losses["types"][0]], "coef": [losses["coef"][0]]}
    model.eval()
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = model.device
    indices = list(range(len(eval_dataset)))[rank::world]
    mdice, channel_dice, loss_all = 0.0, None, 0.0
    if print_detail:
        _log("Start evaluating (total_samples: {}, total_iters: {})...".format(len(eval_dataset), len(indices)))
    with torch.no_grad():
        for it, idx in enumerate(indices):
            im, label = _collate(eval_dataset, [idx], device)
            fused = model.predict_with_losses(im, label, new_loss) if hasattr(model, "predict_with_losses") else None
            if fused is not None:  # 1x1x1 head + argmax + loss sums in one kernel (no logits round trip)
                pred, loss, per_channel_dice = fused
            else:
                tf = getattr(getattr(eval_dataset, "transforms", None), "transforms", None)
                pred, logits = inference(model, im, ori_shape=label.shape[-3:], transforms=tf)
                loss, per_channel_dice = loss_computation([logits], label.to(torch.int32), new_loss)
            loss_all += float(sum(loss))
            mdice += float(np.mean(per_channel_dice))
            channel_dice = per_channel_dice if channel_dice is None else channel_dice + per_channel_dice
            if save_dir is not None and it < 5 and rank == 0:
                os.makedirs(os.path.join(save_dir, str(it)), exist_ok=True)
                np.save(os.path.join(save_dir, str(it), "pred.npy"), pred.cpu().numpy())
    mdice, loss_all, channel_dice = reduce_eval_metrics(mdice, loss_all, channel_dice, len(indices), device,
                                                          getattr(model, "num_classes", 1))
    if print_detail:
        _log("[EVAL] #Images: {}, Dice: {:.4f}, Loss: {:6f}".format(len(eval_dataset), mdice, loss_all))
        _log("[EVAL] Class dice: \n" + str(np.round(channel_dice, 4)))
    return {"mdice": mdice}


def train(model, train_dataset, val_dataset=None, optimizer=None, save_dir="output", iters=10000, batch_size=2,
          resume_model=None, save_interval=1000, log_iters=10, num_workers=0, use_vdl=False, losses=None,
          keep_checkpoint_max=5, profiler_options=None, to_static_training=False, seed=0):
    from .datasets import BatchLoader, DistributedBatchSampler
    model.train()
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = model.device
    start_iter = resume(model, optimizer, resume_model) if resume_model is not None else 0
    os.makedirs(save_dir, exist_ok=True)
    reducer = DistributedGradReducer(model.store.grad).attach(model)
    optimizer.grad_scale = reducer.grad_scale
    # `to_static_training` (the reference switches Paddle to its static-graph mode here, core/train.py:97-99): capture
    # the whole step - at world > 1 including the bucketed NCCL all-reduces - into ONE CUDA graph.  Needs fixed batch
    # shapes; a trailing short batch falls back to the eager step.
    graphed = None
    if to_static_training and reducer.capturable and iters - start_iter >= 2:
        from .graph import GraphedTrainStep
        # the capture's warm-up steps are real optimizer steps on the first batch: never run past `iters`
        graphed = GraphedTrainStep(model, losses, optimizer, reducer=reducer, warmup=min(3, iters - start_iter - 1))
    prof_range = None
    if profiler_options:  # "batch_range=[10,20]" -> cudaProfilerStart/Stop window (ncu / nsys --capture-range)
        import re
        m = re.search(r"batch_range=\[(\d+),\s*(\d+)\]", profiler_options)
        prof_range = (int(m.group(1)), int(m.group(2))) if m else None
    # DistributedBatchSampler(shuffle=True, drop_last=False) + DataLoader(num_workers) (core/train.py:87-95): every rank
    # draws the same number of equally shaped batches (the index list is padded to a multiple of the world size)
    sampler = DistributedBatchSampler(len(train_dataset), batch_size, rank, world, shuffle=True, seed=seed)
    iters_per_epoch = max(len(sampler), 1)
    loader = BatchLoader(train_dataset, iter(sampler), device, num_workers=num_workers)
    avg_loss, mdice, save_models, best_mean_dice, best_model_iter = 0.0, 0.0, deque(), -1.0, -1
    it = start_iter
    n_acc = 0
    reader_cost = batch_cost = 0.0

    def crossed(prev, cur, k):  # a multiple of k lies in (prev, cur]
        return k > 0 and cur // k > prev // k

    batch_start = time.time()
    while it < iters:
        images, labels = next(loader)
        reader_cost += time.time() - batch_start
        prev_it = it
        if prof_range and it == prof_range[0]:
            torch.cuda.cudart().cudaProfilerStart()
        use_graph = graphed is not None and (not graphed.captured or tuple(images.shape) == tuple(graphed.s_img.shape))
        if use_graph:
            first = not graphed.captured
            lr = optimizer.get_lr()
            loss, per_channel_dice = graphed(images, labels.to(torch.int32))
            it += 1 + (graphed.warmup if first else 0)  # the capture's warm-up steps are real optimizer steps
        else:
            if graphed is not None and graphed.captured:  # odd-shaped batch: eager step with a host-side LR
                saved_lr_dev, optimizer.lr_dev = optimizer.lr_dev, None
                model._defer_prepack = False
            logits_list = model(images)
            loss_list, per_channel_dice = loss_computation(logits_list, labels.to(torch.int32), losses)
            loss = sum(loss_list)
            loss.backward()
            reducer.wait()
            optimizer.step()
            lr = optimizer.get_lr()
            it += 1
            if hasattr(optimizer._learning_rate, "step"):
                optimizer._learning_rate.step()
            model.clear_gradients()
            if graphed is not None and graphed.captured:
                optimizer.lr_dev = saved_lr_dev
                optimizer.lr_dev.fill_(float(optimizer.get_lr()))
                model._defer_prepack = True
                if model._side_stream is not None:  # eager re-pack of this step must not overlap the next replay's
                    torch.cuda.current_stream().wait_stream(model._side_stream)
        if prof_range and prev_it < prof_range[1] <= it:
            torch.cuda.cudart().cudaProfilerStop()
        avg_loss += float(loss.detach())
        mdice += float(np.mean(per_channel_dice)) * 100
        batch_cost += time.time() - batch_start
        n_acc += 1
        if crossed(prev_it, it, log_iters):
            if rank == 0:
                avg_loss /= n_acc  # (== log_iters except after a graph capture, whose warm-up steps are not logged)
                mdice /= n_acc
                bc, rc = batch_cost / n_acc, reader_cost / n_acc
                eta = int((iters - it) * bc)
                _log("[TRAIN] epoch: {}, iter: {}/{}, loss: {:.4f}, DSC: {:.4f}, lr: {:.6f}, batch_cost: {:.4f}, "
                     "reader_cost: {:.5f}, ips: {:.4f} samples/sec | ETA {:02d}:{:02d}:{:02d}".format(
                         it // iters_per_epoch, it, iters, avg_loss, mdice, lr, bc, rc, batch_size / bc,
                         eta // 3600, (eta % 3600) // 60, eta % 60))
            avg_loss = mdice = reader_cost = batch_cost = 0.0
            n_acc = 0
        at_save = crossed(prev_it, it, save_interval) or it >= iters
        if at_save and val_dataset is not None:
            result = evaluate(model, val_dataset, losses, print_detail=True, save_dir=save_dir)
            model.train()
        if at_save and rank == 0:
            cur = os.path.join(save_dir, "iter_{}".format(it))
            save_checkpoint(model, optimizer, cur)
            save_models.append(cur)
            if len(save_models) > keep_checkpoint_max > 0:
                shutil.rmtree(save_models.popleft())
            if val_dataset is not None:
                if result["mdice"] > best_mean_dice:
                    best_mean_dice, best_model_iter = result["mdice"], it
                    save_checkpoint(model, None, os.path.join(save_dir, "best_model"))
                _log("[EVAL] The model with the best validation mDice ({:.4f}) was saved at iter {}.".format(
                    best_mean_dice, best_model_iter))
        batch_start = time.time()
    loader.close()
    time.sleep(0.1)
