"""VNet on hand-written sm_100a CUDA — drop-in for medicalseg/models/vnet.py:178-267 of the reference.

Same constructor signature, attribute tree, state-dict names (Paddle naming: `*.bn1._mean`, `*.bn1._variance`,
`*.relu1._weight`) and `forward(x[N,Cin,D,H,W] f32) -> [logits[N,num_classes,D,H,W] f32]` contract.

Host side = Python over PyTorch tensors (memory, streams, autograd plumbing only); every device op is a kernel of
libmedseg_b200.so called through medicalseg_b200.ops.  Internally activations are blocked-8 ("B8",
[N][C/8][D][H][W][8]) bf16 (compute_dtype='bf16', tensor-core path) or f32 (compute_dtype='f32', CUDA-core parity
path).  Skip-concats are free: producers write straight into channel halves of the pre-allocated xcat buffers.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops
from ..ops import B8

BN_MOMENTUM = 0.9   # paddle.nn.BatchNorm3D default (running = 0.9*running + 0.1*batch)
BN_EPS = 1e-5


def _pad(c: int, m: int) -> int:
    return (c + m - 1) // m * m


# =========================================================================================================
# parameter store: ONE flat f32 buffer for parameters, one for gradients (=> single fused optimizer launch and
# contiguous all-reduce buckets); every parameter is a view.  Physical slots are padded so kernels that work on
# zero-padded channel counts never read out of bounds.
# =========================================================================================================
class _Slot:
    __slots__ = ("name", "shape", "numel", "phys", "offset", "is_buffer", "init", "tap_major")

    def __init__(self, name, shape, phys, is_buffer, init):
        self.name, self.shape, self.is_buffer, self.init = name, tuple(shape), is_buffer, init
        self.numel = int(np.prod(shape))
        self.phys = _pad(max(phys, self.numel), 4)
        self.offset = -1
        # tap_major: a [co][ci][5][5][5] conv weight physically stored as [125][co][ci] - the layout the tensor-core
        # weight-gradient kernels accumulate in (parameters, gradients and momentum share it; `view` permutes back)
        self.tap_major = False


class ParamStore:
    def __init__(self):
        self.slots: "OrderedDict[str, _Slot]" = OrderedDict()
        self.flat = self.grad = self.buffers = None

    def add(self, name, shape, init, phys=0, is_buffer=False):
        self.slots[name] = _Slot(name, shape, phys, is_buffer, init)
        return name

    def finalize(self, device):
        po = bo = 0
        for s in self.slots.values():
            if s.is_buffer:
                s.offset, bo = bo, bo + s.phys
            else:
                s.offset, po = po, po + s.phys
        self.flat = torch.zeros(po, dtype=torch.float32, device=device)
        self.grad = torch.zeros(po, dtype=torch.float32, device=device)
        self.buffers = torch.zeros(max(bo, 4), dtype=torch.float32, device=device)

    def _base(self, s: _Slot):
        return self.buffers if s.is_buffer else self.flat

    def view_of(self, base: torch.Tensor, name) -> torch.Tensor:
        """the slot of `name` inside any flat buffer laid out like the parameters (params, grads, momentum), in the
        reference (Paddle) shape; for tap-major slots this is a permuted, non-contiguous VIEW of the same memory"""
        s = self.slots[name]
        flat = base[s.offset:s.offset + s.numel]
        if s.tap_major:
            co, ci = s.shape[:2]
            return flat.view(-1, co, ci).permute(1, 2, 0).unflatten(2, s.shape[2:])
        return flat.view(s.shape)

    def view(self, name) -> torch.Tensor:
        return self.view_of(self._base(self.slots[name]), name)

    def raw(self, name) -> torch.Tensor:
        """physical 1-D storage of the slot (what the kernels of a tap-major layer read)"""
        s = self.slots[name]
        return self._base(s)[s.offset:s.offset + s.numel]

    def grad_raw(self, name) -> torch.Tensor:
        s = self.slots[name]
        return self.grad[s.offset:s.offset + s.numel]

    def phys(self, name) -> torch.Tensor:
        """physical (zero-padded) 1-D slot, what kernels index"""
        s = self.slots[name]
        return self._base(s)[s.offset:s.offset + s.phys]

    def grad_view(self, name) -> torch.Tensor:
        return self.view_of(self.grad, name)

    def grad_phys(self, name) -> torch.Tensor:
        s = self.slots[name]
        return self.grad[s.offset:s.offset + s.phys]

    def span(self, names: Sequence[str]):
        """[lo, hi) range of the flat parameter buffer covered by `names` (all non-buffers)"""
        sl = [self.slots[n] for n in names if not self.slots[n].is_buffer]
        return min(s.offset for s in sl), max(s.offset + s.phys for s in sl)


class ParamList(list):
    """list of parameter views that also carries the flat buffers (used by optimizer.Momentum / parallel)"""
    store: ParamStore = None


# =========================================================================================================
# layer objects: hold parameter names, run kernels, remember what backward needs
# =========================================================================================================
class _Module:
    """tiny stand-in for nn.Layer so the reference's attribute tree (model.in_tr.conv1 ...) exists"""

    def __init__(self, prefix):
        self._prefix = prefix
        self._param_names: List[str] = []

    def named_children(self):
        for k, v in self.__dict__.items():
            if isinstance(v, _Module):
                yield k, v
            elif isinstance(v, (list, tuple)) and v and isinstance(v[0], _Module):
                for i, m in enumerate(v):
                    yield "%s.%d" % (k, i), m

    def all_param_names(self) -> List[str]:
        out = list(self._param_names)
        for _, m in self.named_children():
            out += m.all_param_names()
        return out


class _BN(_Module):
    def __init__(self, store, prefix, c, c_phys):
        super().__init__(prefix)
        self.c, self.c_phys = c, c_phys
        self.weight = store.add(prefix + ".weight", (c,), ("const", 1.0), c_phys)
        self.bias = store.add(prefix + ".bias", (c,), ("const", 0.0), c_phys)
        self.mean = store.add(prefix + "._mean", (c,), ("const", 0.0), c_phys, is_buffer=True)
        self.var = store.add(prefix + "._variance", (c,), ("const", 1.0), c_phys, is_buffer=True)
        self._param_names = [self.weight, self.bias, self.mean, self.var]


class _PReLU(_Module):
    def __init__(self, store, prefix, c, c_phys):
        super().__init__(prefix)
        self._weight = store.add(prefix + "._weight", (c,), ("const", 0.25), c_phys)
        self._param_names = [self._weight]


class _Conv(_Module):
    def __init__(self, store, prefix, wshape, bias_c, init, bias_phys=0):
        super().__init__(prefix)
        self.weight = store.add(prefix + ".weight", wshape, init)
        self.bias = store.add(prefix + ".bias", (bias_c,), ("const", 0.0), bias_phys)
        self._param_names = [self.weight, self.bias]


class _BnAct:
    """BatchNorm3D(+tile | +residual) + PReLU(s) forward/backward around the msb_bn_* kernels."""

    def __init__(self, eng, bn: _BN, relu1: _PReLU, relu2: Optional[_PReLU] = None):
        self.eng, self.bn, self.relu1, self.relu2 = eng, bn, relu1, relu2
        self.saved = None
        self.bnbuf = None
        self._affine, self._affine_key = None, None

    def eval_affine(self):
        """(scale, shift) of the running-statistics BatchNorm: y*scale + shift with scale = gamma/sqrt(var+eps) and
        shift = beta - mean*scale, cached until the parameters or the running statistics change"""
        eng, st = self.eng, self.eng.store
        key = (eng.param_version, eng.bn_stats_version)
        if self._affine_key != key:
            var = st.phys(self.bn.var).double()
            scale = (st.phys(self.bn.weight).double() / torch.sqrt(var + BN_EPS)).float()  # kernels: f64 rsqrt, f32 product
            shift = st.phys(self.bn.bias) - st.phys(self.bn.mean) * scale
            self._affine, self._affine_key = (scale.contiguous(), shift.contiguous()), key
        return self._affine

    def fwd(self, y: B8, out: B8, sums, residual: Optional[B8] = None, tile=None, tile_c=0):
        eng, st = self.eng, self.eng.store
        g = eng.groups(y.n)
        c = y.c
        if self.bnbuf is None or self.bnbuf.numel() != 4 * g * c:
            self.bnbuf = torch.empty(4 * g * c, dtype=torch.float32, device=y.buf.device)
        count = y.s * (y.n if g == 1 else 1)
        if eng.training and eng.sync_world() > 1:
            eng.stat_all_reduce(sums)  # [2][g][C] f64 sums over all ranks (stream-ordered NCCL call)
            count *= eng.sync_world()
        a2 = st.phys(self.relu2._weight) if (self.relu2 is not None and residual is not None) else None
        ops.bn_fwd_fused(y, out, residual, tile, tile_c, sums if eng.training else None, count,
                         st.phys(self.bn.weight), st.phys(self.bn.bias), st.phys(self.bn.mean), st.phys(self.bn.var),
                         BN_MOMENTUM, BN_EPS, eng.training, self.bnbuf, st.phys(self.relu1._weight), a2, g)
        self.saved = (y, residual, tile, tile_c, count, g)

    def bwd(self, gout: B8, dy: B8, dres: Optional[B8] = None, dres_acc=False):
        eng, st = self.eng, self.eng.store
        y, residual, tile, tile_c, count, g = self.saved
        red = eng.scratch_f64(4 * g * y.c)
        a1 = st.phys(self.relu1._weight)
        a2 = st.phys(self.relu2._weight) if (self.relu2 is not None and residual is not None) else None
        ops.bn_act_bwd_reduce(y, residual, tile, tile_c, gout, self.bnbuf, a1, a2, g, red)
        dg, db = st.grad_phys(self.bn.weight), st.grad_phys(self.bn.bias)
        da1 = st.grad_phys(self.relu1._weight)
        da2 = st.grad_phys(self.relu2._weight) if a2 is not None else None
        if eng.bn_training_bwd and eng.sync_world() > 1:
            # SyncBatchNorm backward: the input gradient needs sum(g1), sum(g1*xhat) over ALL ranks, the parameter
            # gradients stay per-rank sums (the data-parallel all-reduce averages them afterwards)
            local = red.clone()
            eng.stat_all_reduce(red)
            ops.bn_act_bwd_apply(y, residual, tile, tile_c, gout, self.bnbuf, a1, a2, red, count, True, dy, dres,
                                 dres_acc, None, None, None, None, g)
            c = y.c
            lf = local.view(4, c).float()  # g == 1 in batch scope
            db[:c] += lf[0]
            dg[:c] += lf[1]
            da1[:c] += lf[2]
            if da2 is not None:
                da2[:c] += lf[3]
        else:
            ops.bn_act_bwd_apply(y, residual, tile, tile_c, gout, self.bnbuf, a1, a2, red, count, eng.bn_training_bwd,
                                 dy, dres, dres_acc, dg, db, da1, da2, g)
        self.saved = None


class _K5:
    """5x5x5 pad-2 conv: tcgen05 kernels for bf16, direct kernels for the f32 parity path."""

    def __init__(self, eng, conv: _Conv, cin, cout, tap_major=True):
        self.eng, self.conv, self.cin, self.cout = eng, conv, cin, cout
        self.packed_f = self.packed_b = None
        self.packed_version = -1
        self.pack_args = None   # remembered after the first (lazy) pack: lets the engine re-pack ahead of use
        self.pack_event = None  # set when the last pack ran on the engine's side stream
        eng.register_packer(self)
        # bf16 engine: master weight / gradient / momentum of this layer live tap-major ([125][co][ci]), so the weight-
        # gradient atomics land in the gradient buffer itself (no workspace memset, no transposition kernel)
        self.tc3 = bool(getattr(eng, "tc3", False))
        self.tap_major = (bool(tap_major) and eng.dtype == torch.bfloat16) or self.tc3
        eng.store.slots[conv.weight].tap_major = self.tap_major
        self.packed_f_lo = self.packed_b_lo = None

    def repack(self):
        self._pack(*self.pack_args)

    def _pack(self, x_c, out_c):
        eng, st = self.eng, self.eng.store
        if self.packed_version == eng.param_version and self.packed_f is not None:
            eng.wait_pack(self)
            return
        eng.wait_pack(self)  # an older side-stream pack of the same buffers must have finished
        self.pack_args = (x_c, out_c)
        w = st.view(self.conv.weight)
        dev = w.device
        cin_pad, cout_pad = x_c, ops.k5_out_pad(out_c)
        if self.packed_f is None:
            self.packed_f = torch.empty(ops.k5_packed_bytes(cin_pad, cout_pad), dtype=torch.uint8, device=dev)
            # input-gradient operand: reduces over the (padded) output channels, produces the input channels
            self.bk_cin_pad, self.bk_cout_pad = _pad(out_c, 16), ops.k5_out_pad(x_c)
            self.packed_b = torch.empty(ops.k5_packed_bytes(self.bk_cin_pad, self.bk_cout_pad), dtype=torch.uint8,
                                        device=dev)
        if self.tap_major:
            w = st.raw(self.conv.weight)
            # both images from one read of the master weight (msb_conv_k5_pack_tm_pair)
            ops.k5_pack_tm_pair(w, self.packed_f, self.packed_b, self.cout, self.cin, 0, cin_pad, cout_pad,
                                self.bk_cin_pad, self.bk_cout_pad)
            if self.tc3:  # lo parts (w - bf16(w)) of both operand images
                if self.packed_f_lo is None:
                    self.packed_f_lo = torch.empty_like(self.packed_f)
                    self.packed_b_lo = torch.empty_like(self.packed_b)
                ops.k5_pack_tm_pair(w, self.packed_f_lo, self.packed_b_lo, self.cout, self.cin, 1, cin_pad, cout_pad,
                                    self.bk_cin_pad, self.bk_cout_pad)
        else:
            ops.k5_pack(w, self.packed_f, self.cout, self.cin, 0, cin_pad, cout_pad)
            ops.k5_pack(w, self.packed_b, self.cout, self.cin, 1, self.bk_cin_pad, self.bk_cout_pad)
        self.packed_version = eng.param_version

    def fwd(self, x: B8, out: B8, sums):
        eng, st = self.eng, self.eng.store
        g = eng.groups(x.n)
        if self.tc3:
            self._pack(x.c, out.c)
            xh, xl = eng.split_hi_lo(x)
            bias = st.view(self.conv.bias)
            ws = eng.splitk_workspace(x.n, out.c, x.dims, x.c)
            ops.k5_fwd(xh, self.packed_f, bias, self.cout, out, False, None, g, None, ws)
            ops.k5_fwd(xl, self.packed_f, None, self.cout, out, True, None, g, None, ws)
            ops.k5_fwd(xh, self.packed_f_lo, None, self.cout, out, True, None, g, sums, ws)
        elif eng.dtype == torch.bfloat16:
            self._pack(x.c, out.c)
            ops.k5_fwd(x, self.packed_f, st.view(self.conv.bias), self.cout, out, False, None, g, sums,
                       eng.splitk_workspace(x.n, out.c, x.dims, x.c))
        else:
            ops.conv_strided_fwd(x, st.view(self.conv.weight), st.view(self.conv.bias), out, (5, 5, 5), (1, 1, 1),
                                 (2, 2, 2), g, sums, self.cin, self.cout)

    def fwd_act(self, x: B8, out: B8, act: "_BnAct", residual: Optional[B8] = None):
        """evaluation-mode LUConv in one kernel: running-statistics BN, PReLU and (block tail) residual + second PReLU
        run in the conv epilogue (msb_conv_k5_fwd_act) instead of a separate pass over the activation"""
        eng, st = self.eng, self.eng.store
        self._pack(x.c, out.c)
        scale, shift = act.eval_affine()
        a2 = st.phys(act.relu2._weight) if residual is not None else None
        ops.k5_fwd_act(x, self.packed_f, st.view(self.conv.bias), self.cout, out, scale, shift,
                       st.phys(act.relu1._weight), residual, a2, eng.splitk_workspace(x.n, out.c, x.dims, x.c))

    def bwd(self, x: B8, dy: B8, dx: Optional[B8], accumulate=False, ch_scale=None):
        eng, st = self.eng, self.eng.store
        w, dw, db = st.view(self.conv.weight), st.grad_view(self.conv.weight), st.grad_view(self.conv.bias)
        if eng.bias_grad_is_zero():
            db = None
        if self.tc3:
            dyh, dyl = eng.split_hi_lo(dy)
            if dx is not None:
                ws = eng.splitk_workspace(dy.n, dx.c, dy.dims, dy.c)
                ops.k5_fwd(dyh, self.packed_b, None, self.cin, dx, accumulate, ch_scale, 1, None, ws)
                ops.k5_fwd(dyl, self.packed_b, None, self.cin, dx, True, ch_scale, 1, None, ws)
                ops.k5_fwd(dyh, self.packed_b_lo, None, self.cin, dx, True, ch_scale, 1, None, ws)
            xh, xl = eng.split_hi_lo(x)
            dw_tm = st.grad_raw(self.conv.weight)
            ops.k5_wgrad_tm(xh, dyh, dw_tm, None, self.cout, self.cin)
            ops.k5_wgrad_tm(xl, dyh, dw_tm, None, self.cout, self.cin)
            ops.k5_wgrad_tm(xh, dyl, dw_tm, None, self.cout, self.cin)
            if db is not None:
                ops.channel_sum(dyh, self.cout, db)  # eval-mode backward only; hi + lo = the f32 gradient
                ops.channel_sum(dyl, self.cout, db)
        elif eng.dtype == torch.bfloat16:
            if dx is not None:
                ops.k5_fwd(dy, self.packed_b, None, self.cin, dx, accumulate, ch_scale, 1, None,
                           eng.splitk_workspace(dy.n, dx.c, dy.dims, dy.c))
            if self.tap_major:
                ops.k5_wgrad_tm(x, dy, st.grad_raw(self.conv.weight), db, self.cout, self.cin)
            else:
                ops.k5_wgrad(x, dy, dw, db, self.cout, self.cin, eng.wgrad_workspace(self.cin, self.cout))
        else:
            if dx is not None:
                if ch_scale is None:
                    ops.conv_strided_bwd_data(dy, w, None, dx, (5, 5, 5), (1, 1, 1), (2, 2, 2), accumulate, 1, None,
                                              self.cout, self.cin)
                else:
                    tmp = B8(dx.n, dx.c, dx.dims, dx.dtype, device=dx.buf.device)
                    ops.conv_strided_bwd_data(dy, w, None, tmp, (5, 5, 5), (1, 1, 1), (2, 2, 2), False, 1, None,
                                              self.cout, self.cin)
                    ops.channel_scale(tmp, dx, ch_scale, accumulate)
            ops.conv_strided_wgrad(x, dy, dw, db, (5, 5, 5), (1, 1, 1), (2, 2, 2), False, self.cin, self.cout)


class _K2S2:
    """tensor-core path of a strided, unpadded down_conv (vnet.py:98-99) or up_conv (vnet.py:133-137) and its input
    gradient: `gather` reduces a window of big-grid voxels into one small-grid voxel, `scatter` expands one into a
    window.  The 5-D weight is [A][B][kd][kh][kw]: gather produces A channels from B, scatter produces B channels from A.
    Handles the default 2x2x2 / stride 2 and the anisotropic MRI kernels (stride 1 along the last axis)."""

    def __init__(self, eng, conv: _Conv, a, b, kernel, stride):
        self.eng, self.conv, self.a, self.b = eng, conv, a, b
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        chan_ok = eng.dtype == torch.bfloat16 and a % 16 == 0 and b % 16 == 0 and a <= 256 and b <= 256
        # packed-operand size 0 = geometry not supported by the tensor-core kernel
        self.ok = [chan_ok and ops.tc_packed_bytes(b, _pad(a, 16), self.kernel, self.stride, 0) > 0,
                   chan_ok and ops.tc_packed_bytes(a, _pad(b, 16), self.kernel, self.stride, 1) > 0]
        self.packed = [None, None]
        self.version = [-1, -1]
        self.pack_args = [None, None]
        self.pack_event = None
        if any(self.ok):
            eng.register_packer(self)

    def repack(self):
        for mode in (0, 1):
            if self.pack_args[mode] is not None:
                self._pack(mode, *self.pack_args[mode])

    def _dims_ok(self, big_dims):
        return all(d >= k and (d - k) % s == 0 for d, k, s in zip(big_dims, self.kernel, self.stride))

    def can_gather(self, big_dims):
        return self.ok[0] and self._dims_ok(big_dims)

    def can_scatter(self, big_dims):
        return self.ok[1] and self._dims_ok(big_dims)

    def _pack(self, mode, c_red_pad, c_out_pad):
        eng = self.eng
        if self.version[mode] == eng.param_version and self.packed[mode] is not None:
            eng.wait_pack(self)
            return self.packed[mode]
        eng.wait_pack(self)
        self.pack_args[mode] = (c_red_pad, c_out_pad)
        w = eng.store.view(self.conv.weight)
        if self.packed[mode] is None:
            self.packed[mode] = torch.empty(ops.tc_packed_bytes(c_red_pad, c_out_pad, self.kernel, self.stride, mode),
                                            dtype=torch.uint8, device=w.device)
        c_red, c_out = (self.b, self.a) if mode == 0 else (self.a, self.b)
        ops.tc_pack(w, self.packed[mode], c_red, c_out, mode, c_red_pad, c_out_pad, self.kernel, self.stride)
        self.version[mode] = eng.param_version
        return self.packed[mode]

    def gather(self, x: B8, out: B8, bias, groups=1, sums=None):
        packed = self._pack(0, x.c, _pad(out.c, 16))
        ops.tc_gather(x, packed, bias, self.a, out, self.kernel, self.stride, groups, sums)

    def scatter(self, x: B8, out: B8, bias, accumulate=False, groups=1, sums=None):
        packed = self._pack(1, x.c, _pad(out.c, 16))
        ops.tc_scatter(x, packed, bias, self.b, out, self.kernel, self.stride, accumulate, groups, sums)


class _K551:
    """w-folded 5x5x1 form of a 5x5x5 conv with <= 3 real channels on one side (bf16 tensor-core path only):
    fold_side 0 = input folded (in_tr.conv1, vnet.py:67-68), 1 = output folded (out_tr.conv1, vnet.py:165-166).
    See include/medseg_b200.h, "w-folded 5x5x1 variant"."""

    def __init__(self, eng, conv: _Conv, cin, cout, fold_side):
        self.eng, self.conv, self.cin, self.cout, self.fold_side = eng, conv, cin, cout, fold_side
        self.cin_f = 5 * cin if fold_side == 0 else cin
        self.cout_f = cout if fold_side == 0 else 5 * cout
        assert self.cin_f <= 16 or fold_side == 1
        assert self.cout_f <= 16 or fold_side == 0
        self.packed_f = self.packed_b = None
        self.packed_version = -1
        self.pack_args = None
        self.pack_event = None
        eng.register_packer(self)

    def repack(self):
        self._pack()

    def _pack(self):
        eng, st = self.eng, self.eng.store
        if self.packed_version == eng.param_version and self.packed_f is not None:
            eng.wait_pack(self)
            return
        eng.wait_pack(self)
        self.pack_args = ()
        w = st.view(self.conv.weight)
        self.f_cin_pad, self.f_cout_pad = _pad(self.cin_f, 16), ops.k5_out_pad(_pad(self.cout_f, 8))
        if self.packed_f is None:
            self.packed_f = torch.empty(ops.k551_packed_bytes(self.f_cin_pad, self.f_cout_pad), dtype=torch.uint8,
                                        device=w.device)
        ops.k551_pack(w, self.packed_f, self.cout, self.cin, 0, self.fold_side, self.f_cin_pad, self.f_cout_pad)
        if self.fold_side == 1:  # input gradient: reduces over the folded outputs, produces the real inputs
            self.b_cin_pad, self.b_cout_pad = _pad(self.cout_f, 16), ops.k5_out_pad(_pad(self.cin_f, 8))
            if self.packed_b is None:
                self.packed_b = torch.empty(ops.k551_packed_bytes(self.b_cin_pad, self.b_cout_pad),
                                            dtype=torch.uint8, device=w.device)
            ops.k551_pack(w, self.packed_b, self.cout, self.cin, 1, self.fold_side, self.b_cin_pad, self.b_cout_pad)
        self.packed_version = eng.param_version

    def fwd(self, x: B8, out: B8, bias, groups=1, sums=None):
        """x: (folded) input view; out: 16-channel folded output (fold_side 1, f32) or the real output (fold_side 0)"""
        self._pack()
        ops.k551_fwd(x, self.packed_f, bias, self.cout_f, out, False, None, groups, sums)

    def bwd_data(self, dp: B8, dx: B8):
        self._pack()
        ops.k551_fwd(dp, self.packed_b, None, self.cin_f, dx, False, None, 1, None)

    def wgrad(self, x: B8, dy: B8):
        st = self.eng.store
        ws = self.eng.workspace(ops.k551_wgrad_workspace_bytes(self.cin, self.cout, self.fold_side))
        ops.k551_wgrad(x, dy, st.grad_view(self.conv.weight), self.cout, self.cin, self.fold_side, ws)


class LUConv(_Module):  # vnet.py:32-43
    def __init__(self, eng, prefix, nchan):
        super().__init__(prefix)
        st = eng.store
        self.relu1 = _PReLU(st, prefix + ".relu1", nchan, nchan)
        self.conv1 = _Conv(st, prefix + ".conv1", (nchan, nchan, 5, 5, 5), nchan, ("conv", nchan * 125))
        self.bn1 = _BN(st, prefix + ".bn1", nchan, nchan)
        self.k5 = _K5(eng, self.conv1, nchan, nchan)
        self.act = None  # built by the owning transition (the last LUConv fuses the residual + relu2)


class InputTransition(_Module):  # vnet.py:57-79
    def __init__(self, eng, prefix, in_channels):
        super().__init__(prefix)
        st = eng.store
        self.num_features, self.in_channels = 16, in_channels
        self.conv1 = _Conv(st, prefix + ".conv1", (16, in_channels, 5, 5, 5), 16, ("conv", in_channels * 125))
        self.bn1 = _BN(st, prefix + ".bn1", 16, 16)
        self.relu1 = _PReLU(st, prefix + ".relu1", 16, 16)
        self.act = _BnAct(eng, self.bn1, self.relu1)
        # 1 input channel (every shipped config): w-folded 5x5x1 tensor-core conv (bf16) / direct kernel (f32).
        # 2, 4, 8, 16 input channels (vnet.py:74-79 tiles x 16/Cin times): the regular 5x5x5 path on a zero-padded view.
        self.k551 = _K551(eng, self.conv1, in_channels, 16, 0) if (eng.dtype == torch.bfloat16 and in_channels == 1) else None
        self.k5 = _K5(eng, self.conv1, in_channels, 16) if in_channels > 1 else None


class DownTransition(_Module):  # vnet.py:82-113
    def __init__(self, eng, prefix, in_ch, n_convs, dropout, stride, kernel):
        super().__init__(prefix)
        st = eng.store
        out_ch = 2 * in_ch
        self.in_ch, self.out_ch, self.if_dropout = in_ch, out_ch, dropout
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        self.down_conv = _Conv(st, prefix + ".down_conv", (out_ch, in_ch, *self.kernel), out_ch,
                               ("conv", in_ch * int(np.prod(self.kernel))))
        self.bn1 = _BN(st, prefix + ".bn1", out_ch, out_ch)
        self.relu1 = _PReLU(st, prefix + ".relu1", out_ch, out_ch)
        self.relu2 = _PReLU(st, prefix + ".relu2", out_ch, out_ch)
        self.ops = [LUConv(eng, "%s.ops.%d" % (prefix, i), out_ch) for i in range(n_convs)]
        self.act_down = _BnAct(eng, self.bn1, self.relu1)
        self.k2 = _K2S2(eng, self.down_conv, out_ch, in_ch, self.kernel, self.stride)
        for i, lu in enumerate(self.ops):
            lu.act = _BnAct(eng, lu.bn1, lu.relu1, self.relu2 if i == n_convs - 1 else None)


class UpTransition(_Module):  # vnet.py:116-156
    def __init__(self, eng, prefix, in_ch, out_ch, n_convs, dropout, dropout2, stride, kernel):
        super().__init__(prefix)
        st = eng.store
        half = out_ch // 2
        self.in_ch, self.out_ch, self.half = in_ch, out_ch, half
        self.if_dropout, self.if_dropout2 = dropout, dropout2
        self.kernel, self.stride = tuple(kernel), tuple(stride)
        self.up_conv = _Conv(st, prefix + ".up_conv", (in_ch, half, *self.kernel), half,
                             ("xavier", in_ch * int(np.prod(self.kernel)), half * int(np.prod(self.kernel))))
        self.bn1 = _BN(st, prefix + ".bn1", half, half)
        self.relu1 = _PReLU(st, prefix + ".relu1", half, half)
        self.relu2 = _PReLU(st, prefix + ".relu2", out_ch, out_ch)
        self.ops = [LUConv(eng, "%s.ops.%d" % (prefix, i), out_ch) for i in range(n_convs)]
        self.act_up = _BnAct(eng, self.bn1, self.relu1)
        self.k2 = _K2S2(eng, self.up_conv, in_ch, half, self.kernel, self.stride)
        for i, lu in enumerate(self.ops):
            lu.act = _BnAct(eng, lu.bn1, lu.relu1, self.relu2 if i == n_convs - 1 else None)


class OutputTransition(_Module):  # vnet.py:159-175
    def __init__(self, eng, prefix, in_channels, num_classes):
        super().__init__(prefix)
        st = eng.store
        # folded path (<= 3 classes, bf16): the 5 kw taps ride in the padding of a 16-channel block and the
        # BN/PReLU/1x1 tail works on a single 8-channel plane
        self.folded = eng.dtype == torch.bfloat16 and num_classes <= 3
        self.c, self.cp = num_classes, (8 if self.folded else _pad(num_classes, 16))
        self.conv1 = _Conv(st, prefix + ".conv1", (num_classes, in_channels, 5, 5, 5), num_classes,
                           ("conv", in_channels * 125), self.cp)
        self.bn1 = _BN(st, prefix + ".bn1", num_classes, self.cp)
        self.conv2 = _Conv(st, prefix + ".conv2", (num_classes, num_classes, 1, 1, 1), num_classes,
                           ("conv", num_classes))
        self.relu1 = _PReLU(st, prefix + ".relu1", num_classes, self.cp)
        self.k5 = _K5(eng, self.conv1, in_channels, num_classes, tap_major=not self.folded)  # folded: Paddle layout
        self.k551 = _K551(eng, self.conv1, in_channels, num_classes, 1) if self.folded else None
        self.act = _BnAct(eng, self.bn1, self.relu1)


class _VNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, engine):
        ctx.engine = engine
        return engine._forward(x, record=True)

    @staticmethod
    def backward(ctx, grad_logits):
        ctx.engine._backward(grad_logits)
        return None, None, None


class VNet(_Module):
    """B200-native VNet (reference: medicalseg/models/vnet.py:178-267).

    Extra keyword arguments (not in the reference): `compute_dtype` ('bf16' tensor-core path | 'f32' parity path),
    `stat_scope` ('batch' = reference BatchNorm semantics | 'instance'), `device`.
    """

    _OUT_PREFIX = "out_tr"  # state-dict prefix of the output transition (VNetDeepSup: "out_tr32")

    # hooks of the deep-supervision variant (models/vnet_deepsup.py); no-ops for the plain VNet
    def _build_aux_heads(self):
        pass

    def _after_forward(self, tape):
        pass

    def _aux_dgrad(self, key, g_buf):
        pass

    def __init__(self, elu=False, in_channels=1, num_classes=4, pretrained=None,
                 kernel_size=((2, 2, 2), (2, 2, 2), (2, 2, 2), (2, 2, 2)),
                 stride_size=((2, 2, 2), (2, 2, 2), (2, 2, 2), (2, 2, 2)),
                 compute_dtype="bf16", stat_scope="batch", device=None, seed=None, sync_bn=False):
        super().__init__("")
        if elu:
            raise NotImplementedError("elu=True (nn.ELU) is not supported; the reference itself reports NaN gradients "
                                      "with it (medicalseg/core/train.py:139)")
        if in_channels not in (1, 2, 4, 8, 16):
            # vnet.py:78-79: x is tiled int(16 / in_channels) times and added to the 16-channel conv output
            raise ValueError("in_channels must divide 16 (InputTransition tiles the input to 16 channels), got %r"
                             % (in_channels,))
        if not torch.cuda.is_available():
            raise RuntimeError("medicalseg_b200.VNet needs a CUDA device (sm_100a); there is no CPU fallback")
        from .. import _lib
        _lib.load()  # fail loudly if the CUDA library is missing
        self.best_loss = 1000000
        self.num_classes, self.in_channels, self.pretrained = num_classes, in_channels, pretrained
        self.device = torch.device(device or ("cuda:%d" % torch.cuda.current_device()))
        # 'bf16': tensor-core path; 'f32': CUDA-core parity path; 'f32x3': f32 storage with the 5x5x5 convs on tensor
        # cores as three bf16 passes (hi*hi + lo*hi + hi*lo, ~2^-16 relative) - BASELINE configs[2]
        self.dtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f32x3": torch.float32}[compute_dtype]
        self.tc3 = compute_dtype == "f32x3"
        self.stat_scope = stat_scope
        # sync_bn=True: batch statistics over ALL ranks, the reference's behaviour at world > 1 (cvlibs/config.py:322
        # converts every BatchNorm to SyncBatchNorm).  Costs two tiny f64 all-reduces per BN layer (forward sums,
        # backward sums); the default keeps statistics per rank (north_star: "allreduce for the gradient step only").
        self.sync_bn = bool(sync_bn)
        self.training = True
        self.bn_training_bwd = True
        self.bn_stats_version = 0
        self.param_version = 0
        self.store = ParamStore()
        k, s = [tuple(v) for v in kernel_size], [tuple(v) for v in stride_size]
        self.kernel_size, self.stride_size = k, s
        self.in_tr = InputTransition(self, "in_tr", in_channels)
        self.down_tr32 = DownTransition(self, "down_tr32", 16, 1, False, s[0], k[0])
        self.down_tr64 = DownTransition(self, "down_tr64", 32, 2, False, s[1], k[1])
        self.down_tr128 = DownTransition(self, "down_tr128", 64, 3, True, s[2], k[2])
        self.down_tr256 = DownTransition(self, "down_tr256", 128, 2, True, s[3], k[3])
        self.up_tr256 = UpTransition(self, "up_tr256", 256, 256, 2, True, True, s[3], k[3])
        self.up_tr128 = UpTransition(self, "up_tr128", 256, 128, 2, True, True, s[2], k[2])
        self.up_tr64 = UpTransition(self, "up_tr64", 128, 64, 1, False, False, s[1], k[1])
        self.up_tr32 = UpTransition(self, "up_tr32", 64, 32, 1, False, False, s[0], k[0])
        self.out_tr = OutputTransition(self, self._OUT_PREFIX, 32, num_classes)
        self._build_aux_heads()  # VNetDeepSup: extra parameter slots AFTER the main head (reference order)
        self.store.finalize(self.device)
        self._init_parameters(seed)
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._scratch = torch.zeros(1 << 18, dtype=torch.float64, device=self.device)
        self._scratch_off = 0
        self._wg_ws = None
        self._k2_ws = None
        self._sk_ws = None
        self._side_stream = None
        self._defer_prepack = False  # GraphedTrainStep: the re-pack is issued at the START of the captured step
        self._masks: Optional[Dict[str, torch.Tensor]] = None
        self._drawn: Optional[Dict[str, torch.Tensor]] = None
        self._dropout_step = None
        self._dropout_seed = int(torch.initial_seed() if seed is None else seed) * 2654435761 + 12345
        self._tape = None
        self.grad_ready_hook = None  # callable(lo, hi) on flat-grad ranges, fired in backward order (DDP buckets)
        self.stat_all_reduce = self._torch_stat_all_reduce  # replaced by DistributedGradReducer.attach (own NCCL comm)
        if pretrained is not None:
            self.init_weight()

    # ---------------------------------------------------------------- nn.Layer-like API
    def init_weight(self):
        if self.pretrained is not None:
            from ..utils import load_entire_model
            load_entire_model(self, self.pretrained)

    def train(self):
        self.training = True
        return self

    def eval(self):
        self.training = False
        return self

    def parameters(self) -> ParamList:
        pl = ParamList()
        for name, s in self.store.slots.items():
            if not s.is_buffer:
                p = self.store.view(name)
                p.grad = self.store.grad_view(name)
                pl.append(p)
        pl.store = self.store
        pl.owner = self
        return pl

    def named_parameters(self):
        for name, s in self.store.slots.items():
            if not s.is_buffer:
                yield name, self.store.view(name)

    def state_dict(self):
        return OrderedDict((name, self.store.view(name).detach().contiguous().clone()) for name in self.store.slots)

    def set_state_dict(self, sd, strict=True):
        missing = []
        for name, s in self.store.slots.items():
            if name not in sd:
                missing.append(name)
                continue
            v = torch.as_tensor(np.asarray(sd[name]) if not torch.is_tensor(sd[name]) else sd[name])
            if tuple(v.shape) != s.shape:
                raise ValueError("shape mismatch for %s: %s vs %s" % (name, tuple(v.shape), s.shape))
            self.store.view(name).copy_(v.to(self.device, torch.float32))
        if strict and missing:
            raise KeyError("missing keys in state dict: %s" % missing[:5])
        self.param_version += 1
        return missing

    set_dict = set_state_dict
    load_state_dict = set_state_dict

    def clear_gradients(self):
        ops.zero_(self.store.grad)

    clear_grad = clear_gradients

    def mark_parameters_updated(self):
        """called by the optimizer after a step so packed tensor-core operands are rebuilt"""
        self.param_version += 1
        if not self._defer_prepack:
            self.prepack_async()

    def register_packer(self, packer):
        if not hasattr(self, "_packers"):
            self._packers = []
        self._packers.append(packer)

    def prepack_async(self):
        """Re-packs the bf16 tensor-core weight images of every layer (in forward order) on a SIDE stream right after
        the optimizer step: the packing kernels are HBM-bound and small, the convolutions they feed are tensor-bound,
        so they overlap instead of sitting on the critical path in front of every conv.  Each layer waits on its own
        event just before its first use (wait_pack)."""
        if self.dtype != torch.bfloat16 or not getattr(self, "async_prepack", True):
            return
        ready = [p for p in self._packers if p.pack_args is not None and (not isinstance(p.pack_args, list) or
                                                                          any(a is not None for a in p.pack_args))]
        if not ready:
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream()
        self._side_stream.wait_stream(cur)
        with torch.cuda.stream(self._side_stream):
            for p in ready:
                p.repack()
                ev = torch.cuda.Event()
                ev.record(self._side_stream)
                p.pack_event = ev

    def wait_pack(self, packer):
        ev = packer.pack_event
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            packer.pack_event = None

    def set_dropout_masks(self, masks: Optional[Dict[str, torch.Tensor]], persistent: bool = False):
        """explicit Dropout3D masks ([N,C] of 0 / 2) for the next train-mode forward (every forward until changed when
        `persistent`); None -> draw internally"""
        self._masks = None if masks is None else {k: v.to(self.device, torch.float32).contiguous()
                                                  for k, v in masks.items()}
        self._masks_persistent = bool(persistent) and masks is not None

    def __call__(self, x):
        return self.forward(x)

    def forward(self, x):  # vnet.py:256-267 -> [logits]
        if not x.is_cuda:
            raise RuntimeError("VNet.forward needs a CUDA tensor (no CPU fallback)")
        x = x.to(torch.float32).contiguous()
        if torch.is_grad_enabled():
            return [_VNetFunction.apply(x, self._anchor, self)]
        return [self._forward(x, record=False)]

    def predict_with_losses(self, x, labels=None, losses=None):
        """Evaluation fast path (core/val.py:101-118 = infer.inference argmax + loss_computation of the same logits):
        out_tr's 1x1x1 conv, the argmax and the Dice / CE sums run in ONE kernel, so the logits never reach HBM.
        Returns (pred int32 [N,1,D,H,W], loss_list, per_channel_dice) - the last two None without labels - or None when
        the loss configuration is not one the fused head covers (the caller then uses forward() + loss_computation)."""
        from . import losses as L
        if self.training:
            raise RuntimeError("predict_with_losses is an eval-mode path: call model.eval() first")
        if not x.is_cuda:
            raise RuntimeError("VNet.predict_with_losses needs a CUDA tensor (no CPU fallback)")
        plan = L.fused_head_plan(losses) if labels is not None else ()
        if plan is None:
            return None
        x = x.to(torch.float32).contiguous()
        with torch.no_grad():
            ao = self._forward(x, record=False, head=False)
            ot, st = self.out_tr, self.store
            return L.fused_head_losses(ao, st.view(ot.conv2.weight), st.view(ot.conv2.bias), self.num_classes,
                                       tuple(x.shape[2:]), labels, losses, plan)

    # ---------------------------------------------------------------- helpers
    def groups(self, n):
        return 1 if self.stat_scope == "batch" else n

    def sync_world(self):
        """number of ranks the BatchNorm statistics are shared with (1 = per-rank statistics)"""
        if not self.sync_bn or self.stat_scope != "batch":
            return 1
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    @staticmethod
    def _torch_stat_all_reduce(t):
        torch.distributed.all_reduce(t)

    def scratch_f64(self, count):
        count = _pad(count, 2)
        if self._scratch_off + count > self._scratch.numel():
            raise RuntimeError("f64 scratch pool exhausted")
        v = self._scratch[self._scratch_off:self._scratch_off + count]
        self._scratch_off += count
        return v

    def bias_grad_is_zero(self):
        """A conv bias that feeds a BatchNorm using batch (or instance) statistics has an identically zero gradient:
        the normalisation subtracts the per-channel mean, so sum_v dL/dy[v, c] = 0 exactly.  The bf16 path leaves the
        (pre-zeroed) gradient slot untouched instead of summing bf16 rounding noise over millions of voxels; the f32
        parity path still computes the sum like the reference's autograd does."""
        return (self.dtype == torch.bfloat16 or self.tc3) and self.bn_training_bwd

    def strided_wgrad(self, big: B8, small: B8, dw, dbias, kernel, stride, bias_from_big):
        """weight gradient of a down / up conv: tensor-core path (pointwise GEMM over the tap sub-lattices) for bf16"""
        if self.bias_grad_is_zero():
            dbias = None
        kernel, stride = tuple(kernel), tuple(stride)
        taps = kernel[0] * kernel[1] * kernel[2]
        tc_ok = (big.c in (16, 32, 64, 128) and (taps * big.c) % 128 == 0 and small.c % 16 == 0 and small.c <= 256
                 and max(stride[1], stride[2]) <= 4
                 and all(d >= k and (d - k) % s == 0 for d, k, s in zip(big.dims, kernel, stride)))
        if tc_ok and (self.dtype == torch.bfloat16 or self.tc3):
            need = ops.tc_wgrad_workspace_bytes(big.c, small.c, kernel)
            if self._k2_ws is None or self._k2_ws.numel() < need:
                self._k2_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            if self.tc3:  # f32 storage: three bf16 passes (hi*hi + lo*hi + hi*lo) accumulate into the same f32 dw
                bh, bl = self.split_hi_lo(big)
                sh, sl = self.split_hi_lo(small)
                ops.tc_wgrad(bh, sh, dw, None, kernel, stride, bias_from_big, self._k2_ws)
                ops.tc_wgrad(bl, sh, dw, None, kernel, stride, bias_from_big, self._k2_ws)
                ops.tc_wgrad(bh, sl, dw, None, kernel, stride, bias_from_big, self._k2_ws)
                if dbias is not None:  # eval-mode backward only
                    sh2, sl2 = (bh, bl) if bias_from_big else (sh, sl)
                    ops.channel_sum(sh2, dbias.numel(), dbias)
                    ops.channel_sum(sl2, dbias.numel(), dbias)
            else:
                ops.tc_wgrad(big, small, dw, dbias, kernel, stride, bias_from_big, self._k2_ws)
        else:
            ops.conv_strided_wgrad(big, small, dw, dbias, kernel, stride, (0, 0, 0), bias_from_big)

    def splitk_workspace(self, n, cout_view, dims, cin_view):
        """scratch for the split-K 5x5x5 conv on small volumes (None when the shape does not use it): one private
        partial-sum copy per K slice, no initial contents required"""
        need = ops.k5_fwd_workspace_bytes(n, cout_view, dims, cin_view)
        if need == 0:
            return None
        if self._sk_ws is None or self._sk_ws.numel() < need:
            self._sk_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._sk_ws

    def split_hi_lo(self, x: B8):
        """f32 B8 activation -> (hi, lo) bf16 B8 tensors for the 3 x bf16 tensor-core passes"""
        hi = B8(x.n, x.c, x.dims, torch.bfloat16, device=self.device)
        lo = B8(x.n, x.c, x.dims, torch.bfloat16, device=self.device)
        ops.split_hi_lo(x, hi, lo)
        return hi, lo

    def workspace(self, nbytes):
        if self._wg_ws is None or self._wg_ws.numel() < nbytes:
            self._wg_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._wg_ws

    def wgrad_workspace(self, cin, cout):
        need = ops.k5_wgrad_workspace_bytes(cin, cout)
        if self._wg_ws is None or self._wg_ws.numel() < need:
            self._wg_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._wg_ws

    def _init_parameters(self, seed):
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        for name, s in self.store.slots.items():
            kind = s.init[0]
            if kind == "const":
                v = torch.full(s.shape, float(s.init[1]))
            elif kind == "conv":   # Paddle Conv3D default: Normal(0, sqrt(2/fan_in))
                v = torch.randn(s.shape, generator=g) * math.sqrt(2.0 / s.init[1])
            else:                   # Paddle Conv3DTranspose default: XavierUniform
                bound = math.sqrt(6.0 / (s.init[1] + s.init[2]))
                v = (torch.rand(s.shape, generator=g) * 2 - 1) * bound
            self.store.view(name).copy_(v.to(self.device))

    def _new(self, n, c, dims, zero=False):
        return B8(n, c, dims, self.dtype, device=self.device, zero=zero)

    def _mask(self, site, n, c):
        if not self.training:
            return None
        if self._masks is not None:
            return self._masks[site]
        return self._drawn[site]

    _DROPOUT_SITES = (("down_tr128", 128), ("down_tr256", 256), ("up_tr256.x", 256), ("up_tr256.skip", 128),
                      ("up_tr128.x", 256), ("up_tr128.skip", 64))  # vnet.py:103,108,144-145,149-150 via :201-232

    def _draw_masks(self, n):
        """Dropout3D(p=0.5) masks of all six sites in ONE kernel launch (msb_dropout_masks): counter-based generator
        keyed by (seed, device-side step counter), so eager steps and CUDA-graph replays both draw fresh masks"""
        if self._dropout_step is None:
            self._dropout_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        total = sum(c for _, c in self._DROPOUT_SITES)
        buf = torch.empty(n * total, dtype=torch.float32, device=self.device)
        ops.dropout_masks(self._dropout_seed, self._dropout_step, buf, 0.5)
        out, off = {}, 0
        for site, c in self._DROPOUT_SITES:
            out[site] = buf[off:off + n * c].view(n, c)
            off += n * c
        return out

    @staticmethod
    def _down_dims(dims, k, s):
        return tuple((d - kk) // ss + 1 for d, kk, ss in zip(dims, k, s))

    # ---------------------------------------------------------------- forward
    def _forward(self, x: torch.Tensor, record: bool = False, head: bool = True):
        st, T = self.store, self.training
        n = x.shape[0]
        g = self.groups(n)
        ops.zero_(self._scratch)
        self._scratch_off = 0
        tape = {"x": x, "n": n}
        dims = [tuple(x.shape[2:])]
        for lvl in range(4):
            dims.append(self._down_dims(dims[-1], self.kernel_size[lvl], self.stride_size[lvl]))
        tape["dims"] = dims

        def sums(c):
            return self.scratch_f64(2 * g * c) if T else None

        # evaluation (running-statistics BN, nothing recorded for backward): every LUConv is ONE kernel - BN, PReLU and
        # the block's residual tail run in the conv epilogue (eval_fused_epilogue=False keeps the separate BN pass)
        fuse_eval = (not T and not record and self.dtype == torch.bfloat16 and not self.tc3
                     and getattr(self, "eval_fused_epilogue", True))
        if T:
            self.bn_stats_version += 1  # a train-mode forward moves the running statistics
            if self._masks is None:
                self._drawn = self._draw_masks(n)

        # concat buffers: [up-branch | skip] (vnet.py:152 order)
        xcat32 = self._new(n, 32, dims[0])
        xcat64 = self._new(n, 64, dims[1])
        xcat128 = self._new(n, 128, dims[2])
        xcat256 = self._new(n, 256, dims[3])

        # ---- in_tr
        it = self.in_tr
        y0 = self._new(n, 16, dims[0])
        s0 = sums(16)
        if it.k551 is not None:
            xf = self._new(n, 16, dims[0])
            ops.fold_w_f32(x, self.in_channels, xf, 1)
            it.k551.fwd(xf, y0, st.view(it.conv1.bias), g, s0)
            tape["xf"] = xf
        elif it.k5 is not None:
            xb = B8.from_ncdhw(x, self.dtype, c_pad=16 if (self.dtype == torch.bfloat16 or self.tc3) else
                               _pad(self.in_channels, 8))
            it.k5.fwd(xb, y0, s0)
            tape["xb"] = xb
        else:
            ops.conv_in_fwd(x, st.view(it.conv1.weight), st.view(it.conv1.bias), y0, g, s0)
        out16 = xcat32.view(16, 16)
        it.act.fwd(y0, out16, s0, tile=x, tile_c=self.in_channels)

        # ---- encoder
        def down(tr: DownTransition, xin: B8, lvl: int, out: B8, site: str):
            c = tr.out_ch
            yd = self._new(n, c, dims[lvl])
            sd = sums(c)
            if tr.k2.can_gather(xin.dims):
                tr.k2.gather(xin, yd, st.view(tr.down_conv.bias), g, sd)
            else:
                ops.conv_strided_fwd(xin, st.view(tr.down_conv.weight), st.view(tr.down_conv.bias), yd, tr.kernel,
                                     tr.stride, (0, 0, 0), g, sd)
            dwn = self._new(n, c, dims[lvl])
            tr.act_down.fwd(yd, dwn, sd)
            mask = self._mask(site, n, c) if tr.if_dropout else None
            cur = dwn
            if mask is not None:
                cur = self._new(n, c, dims[lvl])
                ops.channel_scale(dwn, cur, mask, False)
            rec = {"xin": xin, "down": dwn, "mask": mask, "lu_in": []}
            for i, lu in enumerate(tr.ops):
                rec["lu_in"].append(cur)
                last = i == len(tr.ops) - 1
                nxt = out if last else self._new(n, c, dims[lvl])
                if fuse_eval:
                    lu.k5.fwd_act(cur, nxt, lu.act, residual=dwn if last else None)
                else:
                    y = self._new(n, c, dims[lvl])
                    sl = sums(c)
                    lu.k5.fwd(cur, y, sl)
                    lu.act.fwd(y, nxt, sl, residual=dwn if last else None)
                cur = nxt
            return rec

        out32 = xcat64.view(32, 32)
        out64 = self._new(n, 64, dims[2])
        out128 = self._new(n, 128, dims[3])
        out256 = self._new(n, 256, dims[4])
        tape["d32"] = down(self.down_tr32, out16, 1, out32, "down_tr32")
        tape["d64"] = down(self.down_tr64, out32, 2, out64, "down_tr64")
        tape["d128"] = down(self.down_tr128, out64, 3, out128, "down_tr128")
        tape["d256"] = down(self.down_tr256, out128, 4, out256, "down_tr256")

        # ---- decoder
        def up(tr: UpTransition, xin: B8, skip: Optional[B8], xcat: B8, lvl: int, site: str):
            half, c = tr.half, tr.out_ch
            mx = self._mask(site + ".x", n, tr.in_ch) if tr.if_dropout else None
            ms = self._mask(site + ".skip", n, half) if tr.if_dropout2 else None
            xd = xin
            if mx is not None:
                xd = self._new(n, tr.in_ch, xin.dims)
                ops.channel_scale(xin, xd, mx, False)
            if skip is not None:  # skip not already resident in the right half of xcat
                ops.channel_scale(skip, xcat.view(half, half), ms, False)
            yu = self._new(n, half, dims[lvl])
            su = sums(half)
            if tr.k2.can_scatter(dims[lvl]):
                tr.k2.scatter(xd, yu, st.view(tr.up_conv.bias), False, g, su)
            else:
                ops.conv_strided_bwd_data(xd, st.view(tr.up_conv.weight), st.view(tr.up_conv.bias), yu, tr.kernel,
                                          tr.stride, (0, 0, 0), False, g, su)
            tr.act_up.fwd(yu, xcat.view(0, half), su)
            rec = {"xin": xin, "xd": xd, "mx": mx, "ms": ms, "xcat": xcat, "lu_in": [], "skip_sep": skip is not None}
            cur = xcat
            out = self._new(n, c, dims[lvl])
            for i, lu in enumerate(tr.ops):
                rec["lu_in"].append(cur)
                last = i == len(tr.ops) - 1
                nxt = out if last else self._new(n, c, dims[lvl])
                if fuse_eval:
                    lu.k5.fwd_act(cur, nxt, lu.act, residual=xcat if last else None)
                else:
                    y = self._new(n, c, dims[lvl])
                    sl = sums(c)
                    lu.k5.fwd(cur, y, sl)
                    lu.act.fwd(y, nxt, sl, residual=xcat if last else None)
                cur = nxt
            rec["out"] = out
            return rec

        tape["u256"] = up(self.up_tr256, out256, out128, xcat256, 3, "up_tr256")
        tape["u128"] = up(self.up_tr128, tape["u256"]["out"], out64, xcat128, 2, "up_tr128")
        tape["u64"] = up(self.up_tr64, tape["u128"]["out"], None, xcat64, 1, "up_tr64")
        tape["u32"] = up(self.up_tr32, tape["u64"]["out"], None, xcat32, 0, "up_tr32")

        # ---- out_tr
        ot = self.out_tr
        cp = ot.cp
        yo = self._new(n, cp, dims[0])
        so = sums(cp)
        if ot.folded:
            pf = B8(n, 16, dims[0], torch.float32, device=self.device)
            ot.k551.fwd(tape["u32"]["out"], pf, None)
            ops.unfold_w(pf, st.view(ot.conv1.bias), ot.c, yo, g, so)
        elif fuse_eval:
            ot.k5.fwd_act(tape["u32"]["out"], yo, ot.act)
        else:
            ot.k5.fwd(tape["u32"]["out"], yo, so)
        if fuse_eval and not ot.folded:
            ao = yo
        else:
            ao = self._new(n, cp, dims[0])
            ot.act.fwd(yo, ao, so)
        if not head:  # evaluate(): the fused head consumes the activated features directly (see predict_with_losses)
            self._tape = None
            return ao
        logits = torch.empty((n, self.num_classes, *dims[0]), dtype=torch.float32, device=self.device)
        ops.conv1x1_fwd(ao, st.view(ot.conv2.weight), st.view(ot.conv2.bias), logits, self.num_classes,
                        self.num_classes)
        tape["ao"] = ao
        self._after_forward(tape)
        self._tape = tape if record else None
        if not getattr(self, "_masks_persistent", False):
            self._masks = None
        return logits

    # ---------------------------------------------------------------- backward
    def _fire(self, module: _Module):
        if self.grad_ready_hook is not None:
            lo, hi = self.store.span(module.all_param_names())
            self.grad_ready_hook(lo, hi)

    def _backward(self, dlogits: torch.Tensor):
        tape = self._tape
        if tape is None:
            raise RuntimeError("backward called without a recorded forward")
        st = self.store
        n, dims, x = tape["n"], tape["dims"], tape["x"]
        dlogits = dlogits.contiguous().float()
        self.bn_training_bwd = self.training

        # ---- out_tr
        ot = self.out_tr
        ao = tape["ao"]
        da = self._new(n, ot.cp, dims[0])
        ops.conv1x1_bwd(ao, st.view(ot.conv2.weight), dlogits, da, st.grad_view(ot.conv2.weight),
                        st.grad_view(ot.conv2.bias), self.num_classes, self.num_classes)
        dyo = self._new(n, ot.cp, dims[0])
        ot.act.bwd(da, dyo)
        g_u32 = self._new(n, 32, dims[0])
        if ot.folded:
            dpf = self._new(n, 16, dims[0])
            ops.fold_w(dyo, ot.c, dpf, -1)
            ot.k551.bwd_data(dpf, g_u32)
            ot.k551.wgrad(tape["u32"]["out"], dpf)
            if not self.bias_grad_is_zero():
                ops.channel_sum(dyo, ot.c, st.grad_view(ot.conv1.bias))
        else:
            ot.k5.bwd(tape["u32"]["out"], dyo, g_u32)
        self._fire(ot)

        def lu_chain_bwd(tr, rec, g_out: B8, g_first_in: B8, lvl: int, first_scale):
            """backward through tr.ops; g_first_in receives (+=) the gradient of the first LUConv's input and is
            first written with the residual-branch gradient of the fused (add + relu2) stage."""
            c = tr.out_ch
            g_cur = g_out
            for i in range(len(tr.ops) - 1, -1, -1):
                lu = tr.ops[i]
                last = i == len(tr.ops) - 1
                dy = self._new(n, c, dims[lvl])
                if last:
                    lu.act.bwd(g_cur, dy, dres=g_first_in, dres_acc=False)
                else:
                    lu.act.bwd(g_cur, dy)
                if i == 0:
                    lu.k5.bwd(rec["lu_in"][0], dy, g_first_in, accumulate=True, ch_scale=first_scale)
                else:
                    g_prev = self._new(n, c, dims[lvl])
                    lu.k5.bwd(rec["lu_in"][i], dy, g_prev)
                    g_cur = g_prev

        def up_bwd(tr: UpTransition, rec, g_out: B8, lvl: int, g_skip_sep: Optional[B8]):
            half = tr.half
            g_xcat = self._new(n, tr.out_ch, dims[lvl])
            lu_chain_bwd(tr, rec, g_out, g_xcat, lvl, None)
            # skip half: either stays as a view (no dropout2) or is scaled into the separate skip gradient
            if g_skip_sep is not None:
                ops.channel_scale(g_xcat.view(half, half), g_skip_sep, rec["ms"], False)
            dyu = self._new(n, half, dims[lvl])
            tr.act_up.bwd(g_xcat.view(0, half), dyu)
            g_xin = self._new(n, tr.in_ch, rec["xin"].dims)
            if tr.k2.can_gather(dims[lvl]):
                tr.k2.gather(dyu, g_xin, None)
            else:
                ops.conv_strided_fwd(dyu, st.view(tr.up_conv.weight), None, g_xin, tr.kernel, tr.stride, (0, 0, 0), 1,
                                     None)
            self.strided_wgrad(dyu, rec["xd"], st.grad_view(tr.up_conv.weight), st.grad_view(tr.up_conv.bias),
                               tr.kernel, tr.stride, True)
            if rec["mx"] is not None:
                ops.channel_scale(g_xin, g_xin, rec["mx"], False)
            self._fire(tr)
            return g_xin, g_xcat

        g_u64, g_xcat32 = up_bwd(self.up_tr32, tape["u32"], g_u32, 0, None)
        self._aux_dgrad("u64", g_u64)  # (+= the deep-supervision head's input gradient; no-op for VNet)
        g_out16 = g_xcat32.view(16, 16)
        g_u128, g_xcat64 = up_bwd(self.up_tr64, tape["u64"], g_u64, 1, None)
        self._aux_dgrad("u128", g_u128)
        g_out32 = g_xcat64.view(32, 32)
        g_out64 = self._new(n, 64, dims[2])
        g_u256, _ = up_bwd(self.up_tr128, tape["u128"], g_u128, 2, g_out64)
        self._aux_dgrad("u256", g_u256)
        g_out128 = self._new(n, 128, dims[3])
        g_out256, _ = up_bwd(self.up_tr256, tape["u256"], g_u256, 3, g_out128)

        def down_bwd(tr: DownTransition, rec, g_out: B8, g_xin: B8, lvl: int):
            c = tr.out_ch
            g_down = self._new(n, c, dims[lvl])
            lu_chain_bwd(tr, rec, g_out, g_down, lvl, rec["mask"])
            dyd = self._new(n, c, dims[lvl])
            tr.act_down.bwd(g_down, dyd)
            if tr.k2.can_scatter(g_xin.dims):
                tr.k2.scatter(dyd, g_xin, None, True)
            else:
                ops.conv_strided_bwd_data(dyd, st.view(tr.down_conv.weight), None, g_xin, tr.kernel, tr.stride,
                                          (0, 0, 0), True, 1, None)
            self.strided_wgrad(rec["xin"], dyd, st.grad_view(tr.down_conv.weight), st.grad_view(tr.down_conv.bias),
                               tr.kernel, tr.stride, False)
            self._fire(tr)

        down_bwd(self.down_tr256, tape["d256"], g_out256, g_out128, 4)
        down_bwd(self.down_tr128, tape["d128"], g_out128, g_out64, 3)
        down_bwd(self.down_tr64, tape["d64"], g_out64, g_out32, 2)
        down_bwd(self.down_tr32, tape["d32"], g_out32, g_out16, 1)

        # ---- in_tr (no input gradient: the image has stop_gradient=True, core/train.py:123)
        it = self.in_tr
        dy0 = self._new(n, 16, dims[0])
        it.act.bwd(g_out16, dy0)
        if it.k551 is not None:
            it.k551.wgrad(tape["xf"], dy0)
            if not self.bias_grad_is_zero():
                ops.channel_sum(dy0, 16, st.grad_view(it.conv1.bias))
        elif it.k5 is not None:
            it.k5.bwd(tape["xb"], dy0, None)
        else:
            ops.conv_in_wgrad(x, dy0, st.grad_view(it.conv1.weight), st.grad_view(it.conv1.bias))
        self._fire(it)
        self._tape = None

    # ---------------------------------------------------------------- reference self-check (vnet.py:269-282)
    def test(self):
        np.random.seed(1)
        a = np.random.rand(1, self.in_channels, 32, 32, 32)
        x = torch.tensor(a, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            out = self.forward(x)[0]
        print("out", float(out.mean()), float(x.mean()))
        assert tuple(out.shape) == (1, self.num_classes, 32, 32, 32)
        print("Vnet test is complete")
