"""ctypes binding of libmedseg_b200.so (include/medseg_b200.h).  The product path has NO fallback:
if the CUDA library is missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmedseg_b200.so")

MSB_F32, MSB_BF16 = 0, 1


class MsbTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n_stride", C.c_int64), ("c", C.c_int32), ("dtype", C.c_int32)]


class MsbDim3(C.Structure):
    _fields_ = [("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32)]


P, I, L, F, D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
T, D3, SZ = MsbTensor, MsbDim3, C.c_size_t

# name -> (restype, argtypes); mirrors include/medseg_b200.h one to one
SIGNATURES = {
    "msb_version": (I, []),
    "msb_last_error_string": (C.c_char_p, []),
    "msb_set_tile_scheduler": (I, [I]),
    "msb_zero": (I, [P, SZ, P]),
    "msb_to_blocked": (I, [P, I, I, L, T, P]),
    "msb_from_blocked": (I, [T, P, I, I, L, P]),
    "msb_bn_stats": (I, [T, I, L, I, P, P]),
    "msb_bn_finalize": (I, [P, D, P, P, P, P, F, F, I, I, I, P, P]),
    "msb_bn_act_fwd": (I, [T, T, T, P, I, P, P, P, I, L, I, P]),
    "msb_bn_fwd_fused": (I, [T, T, T, P, I, P, D, P, P, P, P, F, F, I, P, P, P, I, L, I, P]),
    "msb_bn_act_bwd_reduce": (I, [T, T, P, I, T, P, P, P, I, L, I, P, P]),
    "msb_bn_act_bwd_apply": (I, [T, T, P, I, T, P, P, P, P, D, I, T, T, I, P, P, P, P, I, L, I, P]),
    "msb_channel_scale": (I, [T, T, P, I, L, I, P]),
    "msb_conv1x1_fwd": (I, [T, P, P, P, I, I, I, L, P]),
    "msb_conv1x1_bwd": (I, [T, P, P, T, P, P, I, I, I, L, P]),
    "msb_conv_in_fwd": (I, [P, P, P, T, I, D3, I, P, P]),
    "msb_conv_in_wgrad": (I, [P, T, P, P, I, D3, P]),
    "msb_conv_strided_fwd": (I, [T, P, P, T, I, D3, D3, D3, D3, I, I, I, P, P]),
    "msb_conv_strided_bwd_data": (I, [T, P, P, T, I, D3, D3, D3, D3, I, I, I, I, P, P]),
    "msb_conv_strided_wgrad": (I, [T, T, P, P, I, D3, D3, D3, D3, I, I, I, P]),
    "msb_conv_tc_packed_bytes": (SZ, [I, I, D3, D3, I]),
    "msb_conv_tc_pack": (I, [P, P, I, I, I, I, I, D3, D3, P]),
    "msb_conv_tc_gather": (I, [T, P, P, I, T, I, D3, D3, D3, I, P, P]),
    "msb_conv_tc_scatter": (I, [T, P, P, I, T, I, D3, D3, D3, I, I, P, P]),
    "msb_conv_tc_wgrad_workspace_bytes": (SZ, [I, I, D3]),
    "msb_conv_tc_wgrad": (I, [T, T, P, P, I, D3, D3, D3, I, P, SZ, P]),
    "msb_conv_k2s2_wgrad_workspace_bytes": (SZ, [I, I, I, D3]),
    "msb_conv_k2s2_wgrad": (I, [T, T, P, P, I, D3, I, P, SZ, P]),
    "msb_conv_k2s2_packed_bytes": (SZ, [I, I]),
    "msb_conv_k2s2_pack": (I, [P, P, I, I, I, I, I, P]),
    "msb_conv_k2s2_gather": (I, [T, P, P, I, T, I, D3, I, P, P]),
    "msb_conv_k2s2_scatter": (I, [T, P, P, I, T, I, D3, I, I, P, P]),
    "msb_conv_k5_packed_bytes": (SZ, [I, I]),
    "msb_conv_k5_out_pad": (I, [I]),
    "msb_conv_k5_pack": (I, [P, P, I, I, I, I, I, P]),
    "msb_conv_k5_fwd": (I, [T, P, P, I, T, I, D3, I, P, I, P, P]),
    "msb_conv_k5_fwd_workspace_bytes": (SZ, [I, I, D3, I]),
    "msb_conv_k5_fwd_ws": (I, [T, P, P, I, T, I, D3, I, P, I, P, P, SZ, P]),
    "msb_conv_k5_fwd_act": (I, [T, P, P, I, T, I, D3, P, P, P, C.POINTER(MsbTensor), P, P, SZ, P]),
    "msb_conv_k5_pack_tm": (I, [P, P, I, I, I, I, I, P]),
    "msb_conv_k5_pack_tm_pair": (I, [P, P, P, I, I, I, I, I, I, I, P]),
    "msb_split_hi_lo": (I, [T, T, T, I, L, P]),
    "msb_conv_k5_wgrad_tm": (I, [T, T, P, P, I, I, I, D3, P]),
    "msb_conv_k5_wgrad_workspace_bytes": (SZ, [I, I]),
    "msb_conv_k5_wgrad": (I, [T, T, P, P, I, I, I, D3, P, SZ, P]),
    "msb_fold_w_f32": (I, [P, I, T, I, D3, I, P]),
    "msb_fold_w": (I, [T, I, T, I, D3, I, P]),
    "msb_unfold_w": (I, [T, P, I, T, I, D3, I, P, P]),
    "msb_conv_k551_packed_bytes": (SZ, [I, I]),
    "msb_conv_k551_pack": (I, [P, P, I, I, I, I, I, I, P]),
    "msb_conv_k551_fwd": (I, [T, P, P, I, T, I, D3, I, P, I, P, P]),
    "msb_conv_k551_wgrad_workspace_bytes": (SZ, [I, I, I]),
    "msb_conv_k551_wgrad": (I, [T, T, P, I, I, I, I, D3, P, SZ, P]),
    "msb_channel_sum": (I, [T, I, I, L, P, P]),
    "msb_debug_set": (I, [I, I]),
    "msb_debug_read_prof": (I, [P]),
    "msb_class_weight_sums": (I, [P, I, I, L, P, P]),
    "msb_class_weight_finalize": (I, [P, D, I, P, P]),
    "msb_dice_ce_fwd": (I, [P, P, P, I, I, L, I, P, P]),
    "msb_dice_ce_finalize": (I, [P, I, P, P]),
    "msb_dice_ce_fwd_ex": (I, [P, P, P, I, I, L, I, I, P, P]),
    "msb_dice_ce_finalize_ex": (I, [P, I, P, P, P]),
    "msb_dice_ce_bwd_ex": (I, [P, P, P, P, I, I, L, I, F, F, P, P, I, P, P]),
    "msb_eval_head": (I, [T, P, P, P, P, I, I, L, I, P, P, P, P]),
    "msb_dice_ce_bwd": (I, [P, P, P, P, I, I, L, I, F, F, P, P, P]),
    "msb_momentum_step": (I, [P, P, P, L, F, F, F, F, P]),
    "msb_momentum_step_lrdev": (I, [P, P, P, L, P, F, F, F, P]),
    "msb_hunorm": (I, [P, P, L, F, F, F, P]),
    "msb_minmax": (I, [P, L, P, P]),
    "msb_normalize": (I, [P, P, L, F, F, P, P]),
    "msb_resample_f32": (I, [P, D3, P, D3, I, I, F, F, F, P]),
    "msb_resample_i32": (I, [P, D3, P, D3, P]),
    "msb_label_remap": (I, [P, L, P, P, I, P]),
    "msb_rotate3d_f32": (I, [P, P, D3, I, I, D, D, D, D, D, D, I, F, P]),
    "msb_rotate3d_i32": (I, [P, P, D3, I, I, D, D, D, D, D, D, I, I, P]),
    "msb_flip3d": (I, [P, P, D3, I, P]),
    "msb_scale_by_max": (I, [P, P, L, P, P]),
    "msb_argmax_channels": (I, [P, I, I, L, P, P]),
    "msb_dropout_masks": (I, [C.c_uint64, P, P, I, F, P]),
    "msb_trilinear_fwd": (I, [P, L, D3, P, D3, P]),
    "msb_trilinear_bwd": (I, [P, L, D3, P, D3, P]),
}

_NO_STATUS = {"msb_version", "msb_last_error_string", "msb_conv_k5_packed_bytes", "msb_conv_k5_out_pad",
              "msb_conv_k2s2_wgrad_workspace_bytes", "msb_conv_k5_fwd_workspace_bytes", "msb_conv_tc_packed_bytes",
              "msb_conv_tc_wgrad_workspace_bytes",
              "msb_conv_k5_wgrad_workspace_bytes", "msb_conv_k551_packed_bytes", "msb_conv_k2s2_packed_bytes",
              "msb_conv_k551_wgrad_workspace_bytes"}

_lock = threading.Lock()
_lib = None


class MsbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the shared library (once).  Raises if it has not been built — there is no CPU fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MsbError(
                "libmedseg_b200.so not found at %s - build it with `python -m medicalseg_b200.build` "
                "(the product path has no CPU/PyTorch fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def call(name: str, *args):
    """Calls an msb_* entry point; non-zero status raises MsbError with msb_last_error_string()."""
    lib = load()
    fn = getattr(lib, name)
    rc = fn(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        msg = lib.msb_last_error_string()
        raise MsbError("%s failed (status %d): %s" % (name, rc, msg.decode() if msg else "?"))
    return rc
