"""Python-side operator layer: torch tensors in, msb_* C-ABI calls out (include/medseg_b200.h).

Every function enqueues on torch's current CUDA stream and never synchronises.  Tensors must live on a
CUDA device: there is no CPU fallback (a CPU tensor raises)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import MSB_BF16, MSB_F32, MsbDim3, MsbTensor, call

_TORCH2MSB = {torch.float32: MSB_F32, torch.bfloat16: MSB_BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.MsbError("medicalseg_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def dim3(d: Sequence[int]) -> MsbDim3:
    return MsbDim3(int(d[0]), int(d[1]), int(d[2]))


class B8:
    """Blocked-8 activation [N][C/8][D][H][W][8]; may be a channel-slice view of a wider buffer."""

    __slots__ = ("buf", "n", "c", "dims", "c_off", "dtype")

    def __init__(self, n: int, c: int, dims: Sequence[int], dtype: torch.dtype, device=None,
                 buf: Optional[torch.Tensor] = None, c_off: int = 0, zero: bool = False):
        assert c % 8 == 0 and c_off % 8 == 0
        self.n, self.c, self.dims, self.c_off, self.dtype = n, c, tuple(int(v) for v in dims), c_off, dtype
        if buf is None:
            alloc = torch.zeros if zero else torch.empty
            buf = alloc((n, c // 8, *self.dims, 8), dtype=dtype, device=device)
        self.buf = buf

    @property
    def s(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]

    @property
    def c_total(self) -> int:
        return self.buf.shape[1] * 8

    def view(self, c_off: int, c: int) -> "B8":
        assert c_off + c <= self.c
        return B8(self.n, c, self.dims, self.dtype, buf=self.buf, c_off=self.c_off + c_off)

    @property
    def mt(self) -> MsbTensor:
        esz = self.buf.element_size()
        ptr = self.buf.data_ptr() + (self.c_off // 8) * self.s * 8 * esz
        return MsbTensor(ptr, self.buf.shape[1] * self.s * 8, self.c, _TORCH2MSB[self.dtype])

    def to_ncdhw(self, c: Optional[int] = None) -> torch.Tensor:
        c = c or self.c
        out = torch.empty((self.n, c, *self.dims), dtype=torch.float32, device=self.buf.device)
        call("msb_from_blocked", self.mt, _ptr(out), self.n, c, self.s, _stream())
        return out

    @staticmethod
    def from_ncdhw(x: torch.Tensor, dtype: torch.dtype, c_pad: Optional[int] = None) -> "B8":
        x = x.contiguous().float()
        n, c = x.shape[:2]
        cp = c_pad or ((c + 7) // 8 * 8)
        out = B8(n, cp, x.shape[2:], dtype, device=x.device)
        call("msb_to_blocked", _ptr(x), n, c, out.s, out.mt, _stream())
        return out


def zero_(t: torch.Tensor) -> torch.Tensor:
    """clears a contiguous CUDA tensor with a stream-ordered memset (no fill kernel)"""
    call("msb_zero", _ptr(t), t.numel() * t.element_size(), _stream())
    return t


NULL_T = MsbTensor(None, 0, 0, 0)


def _mt(t: Optional[B8]) -> MsbTensor:
    return NULL_T if t is None else t.mt


# ---- BatchNorm + PReLU ---------------------------------------------------------------------------------
def bn_stats(x: B8, groups: int, sums: torch.Tensor):
    call("msb_bn_stats", x.mt, x.n, x.s, groups, _ptr(sums), _stream())


def bn_finalize(sums, count, gamma, beta, rmean, rvar, momentum, eps, training, c, groups, bnbuf):
    call("msb_bn_finalize", _ptr(sums), float(count), _ptr(gamma), _ptr(beta), _ptr(rmean), _ptr(rvar),
         float(momentum), float(eps), int(training), c, groups, _ptr(bnbuf), _stream())


def bn_act_fwd(y: B8, out: B8, residual: Optional[B8], tile_src, tile_c, bnbuf, alpha1, alpha2, groups):
    call("msb_bn_act_fwd", y.mt, out.mt, _mt(residual), _ptr(tile_src), int(tile_c), _ptr(bnbuf), _ptr(alpha1),
         _ptr(alpha2), y.n, y.s, groups, _stream())


def bn_fwd_fused(y: B8, out: B8, residual: Optional[B8], tile_src, tile_c, sums, count, gamma, beta, rmean, rvar,
                 momentum, eps, training, bnbuf, alpha1, alpha2, groups):
    call("msb_bn_fwd_fused", y.mt, out.mt, _mt(residual), _ptr(tile_src), int(tile_c), _ptr(sums), float(count),
         _ptr(gamma), _ptr(beta), _ptr(rmean), _ptr(rvar), float(momentum), float(eps), int(training), _ptr(bnbuf),
         _ptr(alpha1), _ptr(alpha2), y.n, y.s, groups, _stream())


def bn_act_bwd_reduce(y: B8, residual, tile_src, tile_c, gout: B8, bnbuf, alpha1, alpha2, groups, red):
    call("msb_bn_act_bwd_reduce", y.mt, _mt(residual), _ptr(tile_src), int(tile_c), gout.mt, _ptr(bnbuf),
         _ptr(alpha1), _ptr(alpha2), y.n, y.s, groups, _ptr(red), _stream())


def bn_act_bwd_apply(y: B8, residual, tile_src, tile_c, gout: B8, bnbuf, alpha1, alpha2, red, count, training,
                     dy: B8, dres: Optional[B8], dres_acc, dgamma, dbeta, dalpha1, dalpha2, groups):
    call("msb_bn_act_bwd_apply", y.mt, _mt(residual), _ptr(tile_src), int(tile_c), gout.mt, _ptr(bnbuf),
         _ptr(alpha1), _ptr(alpha2), _ptr(red), float(count), int(training), dy.mt, _mt(dres), int(dres_acc),
         _ptr(dgamma), _ptr(dbeta), _ptr(dalpha1), _ptr(dalpha2), y.n, y.s, groups, _stream())


def channel_scale(src: B8, dst: B8, scale: Optional[torch.Tensor], accumulate: bool):
    call("msb_channel_scale", src.mt, dst.mt, _ptr(scale), src.n, src.s, int(accumulate), _stream())


# ---- convolutions --------------------------------------------------------------------------------------------
def conv1x1_fwd(a: B8, w, b, logits, ci, co):
    call("msb_conv1x1_fwd", a.mt, _ptr(w), _ptr(b), _ptr(logits), a.n, ci, co, a.s, _stream())


def conv1x1_bwd(a: B8, w, dlogits, da: B8, dw, db, ci, co):
    call("msb_conv1x1_bwd", a.mt, _ptr(w), _ptr(dlogits), da.mt, _ptr(dw), _ptr(db), a.n, ci, co, a.s, _stream())


def conv_in_fwd(x, w, bias, out: B8, groups, sums):
    call("msb_conv_in_fwd", _ptr(x), _ptr(w), _ptr(bias), out.mt, out.n, dim3(out.dims), groups, _ptr(sums),
         _stream())


def conv_in_wgrad(x, dy: B8, dw, dbias):
    call("msb_conv_in_wgrad", _ptr(x), dy.mt, _ptr(dw), _ptr(dbias), dy.n, dim3(dy.dims), _stream())


def conv_strided_fwd(x: B8, w, bias, out: B8, kernel, stride, pad, groups, sums, c_red_real=0, c_out_real=0):
    call("msb_conv_strided_fwd", x.mt, _ptr(w), _ptr(bias), out.mt, x.n, dim3(x.dims), dim3(kernel), dim3(stride),
         dim3(pad), c_red_real, c_out_real, groups, _ptr(sums), _stream())


def conv_strided_bwd_data(x: B8, w, bias, out: B8, kernel, stride, pad, accumulate, groups, sums, c_red_real=0,
                          c_out_real=0):
    call("msb_conv_strided_bwd_data", x.mt, _ptr(w), _ptr(bias), out.mt, x.n, dim3(out.dims), dim3(kernel),
         dim3(stride), dim3(pad), c_red_real, c_out_real, int(accumulate), groups, _ptr(sums), _stream())


def conv_strided_wgrad(big: B8, small: B8, dw, dbias, kernel, stride, pad, bias_from_big, c_big_real=0,
                       c_small_real=0):
    call("msb_conv_strided_wgrad", big.mt, small.mt, _ptr(dw), _ptr(dbias), big.n, dim3(big.dims), dim3(kernel),
         dim3(stride), dim3(pad), c_big_real, c_small_real, int(bias_from_big), _stream())


def k2s2_wgrad_workspace_bytes(n, c_big, c_small, big_dims) -> int:
    return call("msb_conv_k2s2_wgrad_workspace_bytes", n, c_big, c_small, dim3(big_dims))


def k2s2_wgrad(big: B8, small: B8, dw, dbias, bias_from_big, workspace: torch.Tensor):
    call("msb_conv_k2s2_wgrad", big.mt, small.mt, _ptr(dw), _ptr(dbias), big.n, dim3(big.dims), int(bias_from_big),
         _ptr(workspace), workspace.numel() * workspace.element_size(), _stream())


def k2s2_packed_bytes(c_red_pad: int, c_out_pad: int) -> int:
    return call("msb_conv_k2s2_packed_bytes", c_red_pad, c_out_pad)


def k2s2_pack(w, packed, c_red, c_out, mode, c_red_pad, c_out_pad):
    call("msb_conv_k2s2_pack", _ptr(w), _ptr(packed), c_red, c_out, mode, c_red_pad, c_out_pad, _stream())


def k2s2_gather(x: B8, packed, bias, cout, out: B8, groups=1, sums=None):
    call("msb_conv_k2s2_gather", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), groups, _ptr(sums),
         _stream())


def k2s2_scatter(x: B8, packed, bias, cout, out: B8, accumulate=False, groups=1, sums=None):
    call("msb_conv_k2s2_scatter", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(out.dims), int(accumulate),
         groups, _ptr(sums), _stream())


# general strided tensor-core convs (any kernel / stride; see include/medseg_b200.h)
def tc_packed_bytes(c_red_pad, c_out_pad, kernel, stride, mode) -> int:
    return call("msb_conv_tc_packed_bytes", c_red_pad, c_out_pad, dim3(kernel), dim3(stride), mode)


def tc_pack(w, packed, c_red, c_out, mode, c_red_pad, c_out_pad, kernel, stride):
    call("msb_conv_tc_pack", _ptr(w), _ptr(packed), c_red, c_out, mode, c_red_pad, c_out_pad, dim3(kernel),
         dim3(stride), _stream())


def tc_gather(x: B8, packed, bias, cout, out: B8, kernel, stride, groups=1, sums=None):
    call("msb_conv_tc_gather", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), dim3(kernel),
         dim3(stride), groups, _ptr(sums), _stream())


def tc_scatter(x: B8, packed, bias, cout, out: B8, kernel, stride, accumulate=False, groups=1, sums=None):
    call("msb_conv_tc_scatter", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(out.dims), dim3(kernel),
         dim3(stride), int(accumulate), groups, _ptr(sums), _stream())


def tc_wgrad_workspace_bytes(c_big, c_small, kernel) -> int:
    return call("msb_conv_tc_wgrad_workspace_bytes", c_big, c_small, dim3(kernel))


def tc_wgrad(big: B8, small: B8, dw, dbias, kernel, stride, bias_from_big, workspace: torch.Tensor):
    call("msb_conv_tc_wgrad", big.mt, small.mt, _ptr(dw), _ptr(dbias), big.n, dim3(big.dims), dim3(kernel),
         dim3(stride), int(bias_from_big), _ptr(workspace), workspace.numel() * workspace.element_size(), _stream())


def k5_out_pad(c_view: int) -> int:
    return call("msb_conv_k5_out_pad", c_view)


def k5_packed_bytes(cin_pad: int, cout_pad: int) -> int:
    return call("msb_conv_k5_packed_bytes", cin_pad, cout_pad)


def k5_pack(w, packed, cout, cin, mode, cin_pad, cout_pad):
    call("msb_conv_k5_pack", _ptr(w), _ptr(packed), cout, cin, mode, cin_pad, cout_pad, _stream())


def k5_fwd(x: B8, packed, bias, cout, out: B8, accumulate=False, ch_scale=None, groups=1, sums=None,
           workspace: Optional[torch.Tensor] = None):
    """workspace: scratch of >= k5_fwd_workspace_bytes(...) bytes enables the split-K path on small volumes (one private
    partial-sum copy per K slice, summed in fixed order: deterministic); None = regular path only"""
    if workspace is None:
        call("msb_conv_k5_fwd", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), int(accumulate),
             _ptr(ch_scale), groups, _ptr(sums), _stream())
    else:
        call("msb_conv_k5_fwd_ws", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), int(accumulate),
             _ptr(ch_scale), groups, _ptr(sums), _ptr(workspace), workspace.numel() * workspace.element_size(),
             _stream())


def k5_fwd_act(x: B8, packed, bias, cout, out: B8, scale, shift, alpha, residual: Optional[B8] = None, alpha2=None,
               workspace: Optional[torch.Tensor] = None):
    """evaluation-mode LUConv: t = prelu((conv(x) + bias) * scale + shift, alpha); out = prelu(t + residual, alpha2)
    when a residual is given, else t - all in the conv's own epilogue"""
    import ctypes
    res = ctypes.byref(residual.mt) if residual is not None else None
    call("msb_conv_k5_fwd_act", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), _ptr(scale),
         _ptr(shift), _ptr(alpha), res, _ptr(alpha2), _ptr(workspace),
         0 if workspace is None else workspace.numel() * workspace.element_size(), _stream())


def k5_fwd_workspace_bytes(n: int, cout_view: int, dims, cin_view: int) -> int:
    return call("msb_conv_k5_fwd_workspace_bytes", n, cout_view, dim3(dims), cin_view)


def k5_pack_tm(w_tm, packed, cout, cin, mode, cin_pad, cout_pad):
    call("msb_conv_k5_pack_tm", _ptr(w_tm), _ptr(packed), cout, cin, mode, cin_pad, cout_pad, _stream())


def k5_pack_tm_pair(w_tm, packed_f, packed_b, cout, cin, lo_part, f_cin_pad, f_cout_pad, b_cin_pad, b_cout_pad):
    """forward + input-gradient operand images from one read of the tap-major master weight"""
    call("msb_conv_k5_pack_tm_pair", _ptr(w_tm), _ptr(packed_f), _ptr(packed_b), cout, cin, int(lo_part), f_cin_pad,
         f_cout_pad, b_cin_pad, b_cout_pad, _stream())


def split_hi_lo(x: B8, hi: B8, lo: B8):
    call("msb_split_hi_lo", x.mt, hi.mt, lo.mt, x.n, x.s, _stream())


def k5_wgrad_tm(x: B8, dy: B8, dw_tm, dbias, cout, cin):
    call("msb_conv_k5_wgrad_tm", x.mt, dy.mt, _ptr(dw_tm), _ptr(dbias), cout, cin, x.n, dim3(x.dims), _stream())


def k5_wgrad_workspace_bytes(cin: int, cout: int) -> int:
    return call("msb_conv_k5_wgrad_workspace_bytes", cin, cout)


def k5_wgrad(x: B8, dy: B8, dw, dbias, cout, cin, workspace: torch.Tensor):
    call("msb_conv_k5_wgrad", x.mt, dy.mt, _ptr(dw), _ptr(dbias), cout, cin, x.n, dim3(x.dims), _ptr(workspace),
         workspace.numel() * workspace.element_size(), _stream())


# ---- w-folded 5x5x1 convs (in_tr / out_tr) ------------------------------------------------------------------
def fold_w_f32(x: torch.Tensor, c_real: int, out: B8, sign: int):
    call("msb_fold_w_f32", _ptr(x), c_real, out.mt, out.n, dim3(out.dims), sign, _stream())


def fold_w(x: B8, c_real: int, out: B8, sign: int):
    call("msb_fold_w", x.mt, c_real, out.mt, out.n, dim3(out.dims), sign, _stream())


def unfold_w(p: B8, bias, c_real: int, out: B8, groups=1, sums=None):
    call("msb_unfold_w", p.mt, _ptr(bias), c_real, out.mt, out.n, dim3(out.dims), groups, _ptr(sums), _stream())


def k551_packed_bytes(cin_pad: int, cout_pad: int) -> int:
    return call("msb_conv_k551_packed_bytes", cin_pad, cout_pad)


def k551_pack(w, packed, cout, cin, mode, fold_side, cin_pad, cout_pad):
    call("msb_conv_k551_pack", _ptr(w), _ptr(packed), cout, cin, mode, fold_side, cin_pad, cout_pad, _stream())


def k551_fwd(x: B8, packed, bias, cout, out: B8, accumulate=False, ch_scale=None, groups=1, sums=None):
    call("msb_conv_k551_fwd", x.mt, _ptr(packed), _ptr(bias), cout, out.mt, x.n, dim3(x.dims), int(accumulate),
         _ptr(ch_scale), groups, _ptr(sums), _stream())


def k551_wgrad_workspace_bytes(cin: int, cout: int, fold_side: int) -> int:
    return call("msb_conv_k551_wgrad_workspace_bytes", cin, cout, fold_side)


def k551_wgrad(x: B8, dy: B8, dw, cout, cin, fold_side, workspace: torch.Tensor):
    call("msb_conv_k551_wgrad", x.mt, dy.mt, _ptr(dw), cout, cin, fold_side, x.n, dim3(x.dims), _ptr(workspace),
         workspace.numel() * workspace.element_size(), _stream())


def channel_sum(x: B8, c_real: int, out):
    call("msb_channel_sum", x.mt, c_real, x.n, x.s, _ptr(out), _stream())


# ---- loss ---------------------------------------------------------------------------------------------------
def class_weight_sums(logits, psum):
    n, c = logits.shape[:2]
    call("msb_class_weight_sums", _ptr(logits), n, c, logits[0, 0].numel(), _ptr(psum), _stream())


def class_weight_finalize(psum, count, c, weights):
    call("msb_class_weight_finalize", _ptr(psum), float(count), c, _ptr(weights), _stream())


def dice_ce_fwd(logits, labels, class_w, ignore_index, acc, dice_softmax=False):
    n, c = logits.shape[:2]
    call("msb_dice_ce_fwd_ex", _ptr(logits), _ptr(labels), _ptr(class_w), n, c, logits[0, 0].numel(), ignore_index,
         int(bool(dice_softmax)), _ptr(acc), _stream())


def dice_ce_finalize(acc, c, result, dice_w=None):
    call("msb_dice_ce_finalize_ex", _ptr(acc), c, _ptr(dice_w), _ptr(result), _stream())


def eval_head(a: B8, w, b, labels, class_w, c, ignore_index, pred=None, acc=None, psum=None):
    """fused 1x1x1 conv + argmax + Dice/CE sums (+ class-weight softmax sums); the logits never reach HBM"""
    call("msb_eval_head", a.mt, _ptr(w), _ptr(b), _ptr(labels), _ptr(class_w), a.n, c, a.s, ignore_index,
         _ptr(pred), _ptr(acc), _ptr(psum), _stream())


def dice_ce_bwd(logits, labels, class_w, acc, ignore_index, coef_ce, coef_dice, coef_dev, dlogits, dice_w=None,
                dice_softmax=False):
    n, c = logits.shape[:2]
    call("msb_dice_ce_bwd_ex", _ptr(logits), _ptr(labels), _ptr(class_w), _ptr(acc), n, c, logits[0, 0].numel(),
         ignore_index, float(coef_ce), float(coef_dice), _ptr(coef_dev), _ptr(dice_w), int(bool(dice_softmax)),
         _ptr(dlogits), _stream())


def argmax_channels(logits, pred):
    """pred int32 [N,1,D,H,W] = argmax over the channel axis of NCDHW f32 logits (first maximum wins)"""
    n, c = logits.shape[:2]
    call("msb_argmax_channels", _ptr(logits), n, c, logits[0, 0].numel(), _ptr(pred), _stream())


def dropout_masks(seed: int, step_counter, out, p: float = 0.5):
    """out (f32, all Dropout3D sites of one forward concatenated) <- 0 or 1/(1-p); advances the device step counter"""
    call("msb_dropout_masks", int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(step_counter), _ptr(out), out.numel(), float(p),
         _stream())


# ---- optimizer ------------------------------------------------------------------------------------------------
def momentum_step(p, g, v, lr, mu, wd, grad_scale=1.0):
    call("msb_momentum_step", _ptr(p), _ptr(g), _ptr(v), p.numel(), float(lr), float(mu), float(wd),
         float(grad_scale), _stream())


def momentum_step_lrdev(p, g, v, lr_dev, mu, wd, grad_scale=1.0):
    call("msb_momentum_step_lrdev", _ptr(p), _ptr(g), _ptr(v), p.numel(), _ptr(lr_dev), float(mu), float(wd),
         float(grad_scale), _stream())


# ---- preprocessing --------------------------------------------------------------------------------------------
def hunorm(src, dst, hu_min, hu_max, hu_nan):
    call("msb_hunorm", _ptr(src), _ptr(dst), src.numel(), float(hu_min), float(hu_max), float(hu_nan), _stream())


def minmax(src, out2):
    call("msb_minmax", _ptr(src), src.numel(), _ptr(out2), _stream())


def normalize(src, dst, lo, hi, minmax_dev=None):
    call("msb_normalize", _ptr(src), _ptr(dst), src.numel(), float(lo), float(hi), _ptr(minmax_dev), _stream())


def resample_f32(src, dst, order, pre_op=0, p0=0.0, p1=0.0, p2=0.0):
    call("msb_resample_f32", _ptr(src), dim3(src.shape), _ptr(dst), dim3(dst.shape), int(order), int(pre_op),
         float(p0), float(p1), float(p2), _stream())


def resample_i32(src, dst):
    call("msb_resample_i32", _ptr(src), dim3(src.shape), _ptr(dst), dim3(dst.shape), _stream())


def label_remap(labels, keys: Sequence[int], vals: Sequence[int]):
    n = len(keys)
    ka = (C.c_int32 * max(n, 1))(*keys)
    va = (C.c_int32 * max(n, 1))(*vals)
    call("msb_label_remap", _ptr(labels), labels.numel(), C.cast(ka, C.c_void_p), C.cast(va, C.c_void_p), n,
         _stream())


# ---- augmentations ----------------------------------------------------------------------------------------------
def rotate3d(src, dst, axis_a, axis_b, m, off, order, cval=0):
    """m = (m00, m01, m10, m11), off = (off0, off1) as Python floats (f64); src/dst f32 or i32 [D,H,W]"""
    name = "msb_rotate3d_f32" if src.dtype == torch.float32 else "msb_rotate3d_i32"
    cv = float(cval) if src.dtype == torch.float32 else int(cval)
    call(name, _ptr(src), _ptr(dst), dim3(src.shape), int(axis_a), int(axis_b), float(m[0]), float(m[1]), float(m[2]),
         float(m[3]), float(off[0]), float(off[1]), int(order), cv, _stream())


def flip3d(src, dst, axis):
    call("msb_flip3d", _ptr(src), _ptr(dst), dim3(src.shape), int(axis), _stream())


def scale_by_max(src, dst, minmax_dev):
    call("msb_scale_by_max", _ptr(src), _ptr(dst), src.numel(), _ptr(minmax_dev), _stream())


# ---- deep-supervision heads -------------------------------------------------------------------------------------
def trilinear_fwd(src, dst):
    """src [N,C,d,h,w] f32 -> dst [N,C,D,H,W] f32 (F.interpolate trilinear, align_corners=False)"""
    call("msb_trilinear_fwd", _ptr(src), src.shape[0] * src.shape[1], dim3(src.shape[2:]), _ptr(dst),
         dim3(dst.shape[2:]), _stream())


def trilinear_bwd(ddst, dsrc):
    call("msb_trilinear_bwd", _ptr(ddst), ddst.shape[0] * ddst.shape[1], dim3(ddst.shape[2:]), _ptr(dsrc),
         dim3(dsrc.shape[2:]), _stream())
