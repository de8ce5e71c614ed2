"""Direct NCCL binding for the gradient all-reduce (reference: core/train.py:81-88, fleet.distributed_model).

Why not `torch.distributed.all_reduce`: the whole train step is ONE CUDA graph (graph.py), and a collective inside a
capture has to be a plain `ncclAllReduce` enqueued on a stream that belongs to the capture.  ProcessGroupNCCL wraps every
call in its own stream / event / watchdog bookkeeping, which deadlocked when it was captured on 2 x B200 (round 1).  Here
the communicator is ours: `ncclAllReduce(ptr, ptr, count, dtype, sum, comm, stream)` on the reducer's side stream, which
forks from / joins the capturing stream through events - the pattern NCCL documents for CUDA-graph capture.

`torch.distributed` (any backend) is only used to hand the 128-byte ncclUniqueId from rank 0 to the other ranks.
The library is the NCCL that torch itself links (site-packages/nvidia/nccl/lib/libnccl.so.2, 2.28.9).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

NCCL_SUM = 0
_DTYPES = {torch.float32: 7, torch.float64: 8, torch.bfloat16: 9, torch.int32: 2, torch.int64: 4}


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    err = None
    for name in (os.environ.get("MSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"):
        if not name:
            continue
        try:
            lib = C.CDLL(name)
            break
        except OSError as e:  # noqa: PERF203
            err = e
    else:
        raise RuntimeError("libnccl.so.2 not found (%s); the multi-GPU path needs the NCCL torch ships with" % err)
    lib.ncclGetErrorString.restype = C.c_char_p
    lib.ncclGetErrorString.argtypes = [C.c_int]
    lib.ncclGetVersion.argtypes = [C.POINTER(C.c_int)]
    lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    for fn in (lib.ncclGetVersion, lib.ncclGetUniqueId, lib.ncclCommInitRank, lib.ncclAllReduce, lib.ncclCommDestroy):
        fn.restype = C.c_int
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, _load().ncclGetErrorString(rc).decode()))


def version() -> int:
    v = C.c_int()
    _check(_load().ncclGetVersion(C.byref(v)), "ncclGetVersion")
    return v.value


class Communicator:
    """One NCCL communicator over all ranks of the default torch.distributed group (one process per GPU)."""

    def __init__(self, device: torch.device):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("nccl.Communicator needs an initialised torch.distributed group (rendezvous only)")
        lib = _load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        uid = _UniqueId()
        if self.rank == 0:
            _check(lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        payload = [bytes(bytearray(uid.internal)) if self.rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        C.memmove(C.byref(uid), payload[0], 128)
        self.comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")

    def all_reduce_(self, t: torch.Tensor, stream: torch.cuda.Stream = None):
        """in-place sum over ranks, enqueued on `stream` (default: torch's current stream); never synchronises"""
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("nccl all_reduce_ needs a contiguous CUDA tensor")
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _check(_load().ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), _DTYPES[t.dtype], NCCL_SUM, self.comm,
                                     st.cuda_stream), "ncclAllReduce")

    def destroy(self):
        if self.comm:
            _load().ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()
