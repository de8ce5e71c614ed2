"""VNetDeepSup on the B200 engine (reference: medicalseg/models/vnet_deepsup.py:176-281).

The trunk is the VNet engine unchanged.  Added: three `nn.Conv3D(C_stage, num_classes, 3, padding=1)` heads on the
256- / 128- / 64-channel decoder stages (direct-conv kernels: < 0.1 % of the step's FLOPs) whose logits are resized to
the input size with `F.interpolate(mode='trilinear')` (`msb_trilinear_fwd`, adjoint `msb_trilinear_bwd`), the output
transition renamed `out_tr32`, and the parameter slots of `out_tr_all` (built by the reference, never used by its
forward - kept so checkpoints load).  forward -> [out, d1, d2, d3] as the reference returns them.

Backward: the heads' weight gradients depend only on the incoming logits gradients and the recorded decoder features,
so they are computed FIRST (their slots sit at the end of the flat parameter buffer and fire the first data-parallel
bucket); their input gradients are added to the decoder-stage gradients where the trunk's backward produces those
(`VNet._aux_dgrad` hook points)."""
from __future__ import annotations

import torch

from .. import ops
from ..ops import B8
from .vnet import VNet, _BN, _Conv, _Module, _PReLU, _pad


class _AuxHead(_Module):
    """nn.Conv3D(cin, num_classes, kernel_size=3, padding=1), vnet_deepsup.py:245-247.

    bf16 engine: the 3x3x3 kernel is the centre of a zero-bordered 5x5x5 kernel (padding 2 == padding 1 for the
    embedded taps), so forward, input gradient and weight gradient run on the tcgen05 5x5x5 kernels - 4.6x the MACs, but
    at tensor-core rate instead of the CUDA-core direct convolution (measured on the MRI config: the three heads cost
    60 ms per step as direct convolutions).  f32 parity engine: direct kernels."""

    def __init__(self, store, prefix, cin, num_classes):
        super().__init__(prefix)
        self.cin, self.c, self.cp = cin, num_classes, _pad(num_classes, 16)
        self.conv = _Conv(store, prefix, (num_classes, cin, 3, 3, 3), num_classes, ("conv", cin * 27))
        self.w5 = self.dw5 = self.packed_f = self.packed_b = None
        self.version = -1

    def pack(self, eng):
        if self.version == eng.param_version and self.packed_f is not None:
            return
        dev = eng.device
        if self.w5 is None:
            self.w5 = torch.zeros((self.c, self.cin, 5, 5, 5), dtype=torch.float32, device=dev)
            self.dw5 = torch.zeros_like(self.w5)
            self.fw_pad = (self.cin, ops.k5_out_pad(self.cp))             # forward operand: reduce cin -> cp
            self.bw_pad = (self.cp, ops.k5_out_pad(self.cin))             # input-gradient operand: reduce cp -> cin
            self.packed_f = torch.empty(ops.k5_packed_bytes(*self.fw_pad), dtype=torch.uint8, device=dev)
            self.packed_b = torch.empty(ops.k5_packed_bytes(*self.bw_pad), dtype=torch.uint8, device=dev)
        self.w5[:, :, 1:4, 1:4, 1:4].copy_(eng.store.view(self.conv.weight))
        ops.k5_pack(self.w5, self.packed_f, self.c, self.cin, 0, *self.fw_pad)
        ops.k5_pack(self.w5, self.packed_b, self.c, self.cin, 1, *self.bw_pad)
        self.version = eng.param_version

    def fwd(self, eng, x: B8, y: B8):
        st = eng.store
        if eng.dtype == torch.bfloat16:
            self.pack(eng)
            ops.k5_fwd(x, self.packed_f, st.view(self.conv.bias), self.c, y, False, None, 1, None,
                       eng.splitk_workspace(x.n, y.c, x.dims, x.c))
        else:
            ops.conv_strided_fwd(x, st.view(self.conv.weight), st.view(self.conv.bias), y, (3, 3, 3), (1, 1, 1),
                                 (1, 1, 1), 1, None, self.cin, self.c)

    def wgrad(self, eng, x: B8, dy: B8):
        st = eng.store
        dw, db = st.grad_view(self.conv.weight), st.grad_view(self.conv.bias)
        if eng.dtype == torch.bfloat16:
            self.pack(eng)
            self.dw5.zero_()
            ops.k5_wgrad(x, dy, self.dw5, db, self.c, self.cin, eng.wgrad_workspace(self.cin, self.c))
            dw += self.dw5[:, :, 1:4, 1:4, 1:4]
        else:
            ops.conv_strided_wgrad(x, dy, dw, db, (3, 3, 3), (1, 1, 1), (1, 1, 1), False, self.cin, self.c)

    def dgrad_into(self, eng, dy: B8, g_buf: B8):
        if eng.dtype == torch.bfloat16:
            ops.k5_fwd(dy, self.packed_b, None, self.cin, g_buf, True, None, 1, None,
                       eng.splitk_workspace(dy.n, g_buf.c, dy.dims, dy.c))
        else:
            ops.conv_strided_bwd_data(dy, eng.store.view(self.conv.weight), None, g_buf, (3, 3, 3), (1, 1, 1),
                                      (1, 1, 1), True, 1, None, self.c, self.cin)


class _UnusedOutputTransition(_Module):  # out_tr_all, vnet_deepsup.py:248: parameters only (never in the forward)
    def __init__(self, store, prefix, in_channels, num_classes):
        super().__init__(prefix)
        self.conv1 = _Conv(store, prefix + ".conv1", (num_classes, in_channels, 5, 5, 5), num_classes,
                           ("conv", in_channels * 125))
        self.bn1 = _BN(store, prefix + ".bn1", num_classes, num_classes)
        self.conv2 = _Conv(store, prefix + ".conv2", (num_classes, num_classes, 1, 1, 1), num_classes,
                           ("conv", num_classes))
        self.relu1 = _PReLU(store, prefix + ".relu1", num_classes, num_classes)


class _AuxHeads(_Module):
    def __init__(self, store, num_classes):
        super().__init__("")
        self.out_tr64 = _AuxHead(store, "out_tr64", 64, num_classes)
        self.out_tr128 = _AuxHead(store, "out_tr128", 128, num_classes)
        self.out_tr256 = _AuxHead(store, "out_tr256", 256, num_classes)
        self.out_tr_all = _UnusedOutputTransition(store, "out_tr_all", 4 * num_classes, num_classes)


class _VNetDeepSupFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, engine):
        ctx.engine = engine
        main = engine._forward(x, record=True)
        return (main, *engine._extra_logits)

    @staticmethod
    def backward(ctx, g_main, g1, g2, g3):
        eng = ctx.engine
        eng._aux_backward_begin((g1, g2, g3))
        if g_main is None:
            g_main = torch.zeros((eng._tape["n"], eng.num_classes, *eng._tape["dims"][0]), device=eng.device)
        eng._backward(g_main)
        return None, None, None


class VNetDeepSup(VNet):
    """drop-in for medicalseg.models.VNetDeepSup (same constructor arguments, state-dict names and outputs)"""

    _OUT_PREFIX = "out_tr32"
    deep_supervision = True
    _STAGES = (("out_tr256", "u256", 3), ("out_tr128", "u128", 2), ("out_tr64", "u64", 1))  # d1, d2, d3

    def _build_aux_heads(self):
        self.aux = _AuxHeads(self.store, self.num_classes)
        self._extra_logits = []
        self._aux_dy = {}

    @property
    def out_tr32(self):
        return self.out_tr

    @property
    def out_tr64(self):
        return self.aux.out_tr64

    @property
    def out_tr128(self):
        return self.aux.out_tr128

    @property
    def out_tr256(self):
        return self.aux.out_tr256

    @property
    def out_tr_all(self):
        return self.aux.out_tr_all

    def forward(self, x):  # vnet_deepsup.py:256-275 -> [out, d1, d2, d3]
        if not x.is_cuda:
            raise RuntimeError("VNetDeepSup.forward needs a CUDA tensor (no CPU fallback)")
        x = x.to(torch.float32).contiguous()
        if torch.is_grad_enabled():
            return list(_VNetDeepSupFunction.apply(x, self._anchor, self))
        main = self._forward(x, record=False)
        return [main] + self._extra_logits

    # ---- hooks called by the VNet engine ------------------------------------------------------------------------
    def _after_forward(self, tape):
        st, n, dims, c = self.store, tape["n"], tape["dims"], self.num_classes
        outs = []
        for name, key, lvl in self._STAGES:
            head = getattr(self.aux, name)
            y = self._new(n, head.cp, dims[lvl])
            head.fwd(self, tape[key]["out"], y)
            small = y.to_ncdhw(c)
            big = torch.empty((n, c, *dims[0]), dtype=torch.float32, device=self.device)
            ops.trilinear_fwd(small, big)
            outs.append(big)
        self._extra_logits = outs

    def _aux_backward_begin(self, grads):
        tape = self._tape
        if tape is None:
            raise RuntimeError("backward called without a recorded forward")
        st, n, dims, c = self.store, tape["n"], tape["dims"], self.num_classes
        self._aux_dy = {}
        for (name, key, lvl), g in zip(self._STAGES, grads):
            if g is None:
                continue
            head = getattr(self.aux, name)
            small = torch.empty((n, c, *dims[lvl]), dtype=torch.float32, device=self.device)
            ops.trilinear_bwd(g.contiguous().float(), small)
            dy = B8.from_ncdhw(small, self.dtype, c_pad=head.cp)
            head.wgrad(self, tape[key]["out"], dy)
            self._aux_dy[key] = (head, dy)
        self._fire(self.aux)

    def _aux_dgrad(self, key, g_buf):
        item = self._aux_dy.pop(key, None)
        if item is None:
            return
        head, dy = item
        head.dgrad_into(self, dy, g_buf)

    # predict_with_losses (inherited): evaluation scores the main output only (core/val.py:95 keeps the first loss),
    # and its forward stops before the deep-supervision heads
