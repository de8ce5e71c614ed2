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


class DistributedBatchSampler:
    """Index batches of one rank (reference: paddle.io.DistributedBatchSampler(shuffle=True, drop_last=False) as built at
    core/train.py:87-89).  Paddle's contract, kept here: the (shuffled) index list is PADDED with its own leading
    entries to a multiple of the world size, so every rank draws the same number of samples, runs the same number of
    batches and sees the same batch shapes - a data-parallel step can then never wait for a rank whose shard ran dry.
    Ranks take `batch_size`-sized chunks round-robin; the tail (fewer than world*batch_size samples) is split evenly."""

    def __init__(self, num_samples: int, batch_size: int, rank: int = 0, world: int = 1, shuffle: bool = True,
                 seed: int = 0, drop_last: bool = False):
        if num_samples <= 0:
            raise ValueError("DistributedBatchSampler: the dataset is empty")
        if batch_size <= 0 or not 0 <= rank < world:
            raise ValueError("DistributedBatchSampler: bad batch_size / rank / world")
        self.n, self.batch_size, self.rank, self.world = int(num_samples), int(batch_size), int(rank), int(world)
        self.shuffle, self.drop_last = shuffle, drop_last
        self.gen = torch.Generator().manual_seed(seed)
        self.per_rank = (self.n + self.world - 1) // self.world
        self.total = self.per_rank * self.world

    def __len__(self):
        return self.per_rank // self.batch_size if self.drop_last else (self.per_rank + self.batch_size - 1) // self.batch_size

    def epoch(self):
        """list of index lists (one per batch) of this rank for the next epoch"""
        idx = torch.randperm(self.n, generator=self.gen).tolist() if self.shuffle else list(range(self.n))
        while len(idx) < self.total:  # fewer samples than ranks: repeat as often as needed
            idx += idx[:self.total - len(idx)]
        bs, w, r = self.batch_size, self.world, self.rank
        tail = self.total % (bs * w)
        mine = []
        for i in range(r * bs, self.total - tail, bs * w):
            mine += idx[i:i + bs]
        if tail:
            per = tail // w
            rest = idx[self.total - tail:]
            mine += rest[r * per:(r + 1) * per]
        assert len(mine) == self.per_rank
        batches = [mine[i:i + bs] for i in range(0, len(mine), bs)]
        if self.drop_last and batches and len(batches[-1]) < bs:
            batches.pop()
        return batches

    def __iter__(self):
        """endless stream of batches, epoch after epoch (the train loop stops at `iters`)"""
        while True:
            yield from self.epoch()


class BatchLoader:
    """Background, prefetching batch loader (reference: paddle.io.DataLoader(num_workers=..., return_list=True),
    core/train.py:90-95).  `num_workers` threads read samples (np.load releases the GIL) and collate them into a ring of
    PINNED host buffers; each batch is copied host -> device on the worker's own copy stream and handed to the train
    loop together with an event, so `reader_cost` is the time the loop actually blocks - ~0 once the workers keep up
    with a ~12 ms step.  Datasets whose transforms already run on the device (NpyVolumeDataset with a `transforms`
    list) return CUDA tensors; the worker then runs those kernels on its stream and stacks on the device.
    num_workers = 0 loads synchronously on the calling thread (Paddle's meaning of 0)."""

    def __init__(self, dataset, batches, device, num_workers: int = 0, prefetch: int = 2):
        import itertools
        import queue
        import threading
        self.dataset, self.device = dataset, torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_workers = max(int(num_workers), 0)
        self._it = iter(batches)
        self._seq = itertools.count()
        self._lock = threading.Lock()
        self._ready = {}
        self._cv = threading.Condition()
        self._next_out = 0
        self._stop = False
        self._error = None
        self._depth = max(1, prefetch) * max(self.num_workers, 1)
        self._slots = threading.Semaphore(self._depth)
        self._threads = []
        self._queue = queue
        for _ in range(self.num_workers):
            t = threading.Thread(target=self._worker, daemon=True)
            t.start()
            self._threads.append(t)

    # ---- one batch: read + collate + H2D on `stream`; returns (images, labels, event) -------------------------------
    def _load(self, indices, stream, pinned):
        items = [self.dataset[i][:2] for i in indices]
        ims, labs = zip(*items)
        with torch.cuda.stream(stream):
            if ims[0].is_cuda:  # device-side transforms already produced CUDA tensors (on this thread's stream)
                d_im, d_lab = torch.stack(ims), torch.stack(labs)
            else:
                shape_i, shape_l = (len(ims), *ims[0].shape), (len(labs), *labs[0].shape)
                key = (shape_i, ims[0].dtype, shape_l, labs[0].dtype)
                buf = pinned.get(key)
                if buf is None:
                    buf = pinned[key] = [torch.empty(shape_i, dtype=ims[0].dtype, pin_memory=True),
                                         torch.empty(shape_l, dtype=labs[0].dtype, pin_memory=True), None]
                if buf[2] is not None:
                    buf[2].synchronize()  # the previous H2D out of this pinned pair has finished
                torch.stack(ims, out=buf[0])
                torch.stack(labs, out=buf[1])
                d_im = buf[0].to(self.device, non_blocking=True)
                d_lab = buf[1].to(self.device, non_blocking=True)
                buf[2] = torch.cuda.Event()
                buf[2].record(stream)
            ev = torch.cuda.Event()
            ev.record(stream)
        return d_im, d_lab, ev

    def _worker(self):
        try:
            torch.cuda.set_device(self.device)
            stream = torch.cuda.Stream(device=self.device)
            # two pinned pairs per worker alternate, so collating batch k+1 overlaps the H2D copy of batch k
            pinned_sets, use = [{}, {}], 0
            while not self._stop:
                self._slots.acquire()
                if self._stop:
                    return
                with self._lock:
                    try:
                        indices = next(self._it)
                    except StopIteration:
                        indices = None
                    seq = next(self._seq)
                if indices is None:
                    with self._cv:
                        self._ready[seq] = None
                        self._cv.notify_all()
                    return
                out = self._load(indices, stream, pinned_sets[use])
                use ^= 1
                with self._cv:
                    self._ready[seq] = out
                    self._cv.notify_all()
        except BaseException as e:  # surfaced on the consumer side
            with self._cv:
                self._error = e
                self._cv.notify_all()

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.num_workers == 0:
            indices = next(self._it)
            if not hasattr(self, "_sync_pinned"):
                self._sync_pinned = {}
            d_im, d_lab, ev = self._load(indices, cur, self._sync_pinned)
            return d_im, d_lab
        with self._cv:
            while self._next_out not in self._ready and self._error is None:
                self._cv.wait(timeout=1.0)
            if self._error is not None:
                raise self._error
            out = self._ready.pop(self._next_out)
            self._next_out += 1
        self._slots.release()
        if out is None:
            raise StopIteration
        d_im, d_lab, ev = out
        cur.wait_event(ev)
        d_im.record_stream(cur)   # allocated on the worker's stream, consumed on the training stream
        d_lab.record_stream(cur)
        return d_im, d_lab

    def close(self):
        self._stop = True
        for _ in self._threads:
            self._slots.release()
