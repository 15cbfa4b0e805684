"""WaveFlow: host-side mirror of the reference's ``model/waveflow.py`` (``NonCausalLayer2D`` :14-67,
``WN2D`` :70-151, ``WaveFlow`` :154-265) with the same class names, constructor signatures, attribute names and
state-dict keys (``upsampler.1.*``, ``WNs.k.{V,start,layers.i.W,layers.i.W_o}.weight_g/_v``, ``WNs.k.end.weight``,
``invconv1x1.k.weight``).  The modules only own parameters; the arithmetic runs in libcmwg_b200.so:

  WN2D                 -> the WN pipeline with ``cmwg_wn_config.height`` set: 3x3 taps are (line, time) shifts of the
                          same TMA-fed GEMMs, the conditioning is broadcast over lines
  forward (training)   -> per flow ONE autograd node: cmwg_wn_forward(save) + cmwg_waveflow_affine (+ height flip)
                          + cmwg_sum_per_batch; backward = cmwg_waveflow_affine_bwd + cmwg_wn_backward
  reverse (synthesis)  -> per flow ONE call, cmwg_waveflow_inverse_flow: the row-recurrent loop of :243-258 with the
                          reference's rolling per-layer buffers kept as full-height slabs on the device
  upsampler            -> cmwg_upsample_dense_fwd / _bwd
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from . import ops, precision
from .base import FlowBase
from .efficient_modules import InvertibleConv1x1, grad_hint, graph_possible
from .utils import add_weight_norms
from .waveglow import WN, _cond_cache, _SqueezeFunction, _WNState, fused_gate, pack_generation


class NonCausalLayer2D(nn.Module):
    """Parameter container of one 2-D WN layer (reference ``model/waveflow.py:14-67``): ``W`` 3x3 conv Cr -> 2*Cd
    with dilation (h_dilation, dilation), causal in height (top padding h_dilation*(radix-1)), 'same' in time;
    ``W_o`` 1x1 conv Cd -> Cr+Cs (last layer: Cd -> Cs)."""

    def __init__(self, h_dilation, dilation, dilation_channels, residual_channels, skip_channels, radix, bias,
                 last_layer=False):
        super().__init__()
        self.h_pad_size = h_dilation * (radix - 1)
        self.pad_size = dilation * (radix - 1) // 2
        self.W = nn.Conv2d(residual_channels, dilation_channels * 2, kernel_size=radix,
                           dilation=(h_dilation, dilation), bias=bias)
        out_ch = skip_channels if last_layer else residual_channels + skip_channels
        self.W_o = nn.Conv2d(dilation_channels, out_ch, 1, bias=bias)
        self.chs_split = [skip_channels] if last_layer else [residual_channels, skip_channels]

    def forward(self, x, y):
        # Generic per-layer entry for external code that reuses this class; WN2D itself never calls it (its layers
        # run fused in cmwg_wn_forward).
        tmp = F.pad(x, [self.pad_size] * 2 + [self.h_pad_size, 0])
        zw, zf = (self.W(tmp) + y).chunk(2, 1)
        out = self.W_o(fused_gate(zw, zf))
        if len(self.chs_split) == 2:
            res, skip = out.split(self.chs_split, 1)
            return res + x[:, :, -res.size(2):], skip
        return None, out


class WN2D(WN):
    """2-D WaveNet-style transform of WaveFlow (reference ``model/waveflow.py:70-151``): ``start`` 1x1 (1 -> Cr),
    ``V`` conditioning conv for all 8 layers at width rate, 8 ``NonCausalLayer2D`` with width dilation 2^i and the
    height dilations of ``dilation_dict[n_group]``, ``end`` 1x1 (Cs -> 2, no weight norm).  Parameter registration
    order (V, start, layers, end) is the reference's."""

    def __init__(self, n_group, aux_channels, dilation_channels=256, residual_channels=256, skip_channels=256,
                 bias=False, zero_init=True):
        nn.Module.__init__(self)
        dilation_dict = {
            8: [1] * 8,
            16: [1] * 8,
            32: [1, 2, 4] * 2 + [1, 2],
            64: [1, 2, 4, 8, 16, 1, 2, 4],
            128: [1, 2, 4, 8, 16, 32, 64, 1],
        }
        self.h_dilations = dilation_dict[n_group]
        self.dilations = [2 ** i for i in range(8)]
        self.n_group = n_group
        self.in_chs = 1
        self.res_chs = residual_channels
        self.dil_chs = dilation_channels
        self.skp_chs = skip_channels
        self.aux_chs = aux_channels
        self.rdx = 3
        self.has_bias = bool(bias)
        self.r_field = sum(self.dilations) * 2 + 1
        self.h_r_field = sum(self.h_dilations) * 2 + 1

        self.V = nn.Conv1d(aux_channels, dilation_channels * 2 * 8, 1, bias=bias)
        self.V.apply(add_weight_norms)
        self.start = nn.Conv2d(1, residual_channels, 1, bias=bias)
        self.start.apply(add_weight_norms)
        self.layers = nn.ModuleList(
            NonCausalLayer2D(hd, d, dilation_channels, residual_channels, skip_channels, 3, bias, last_layer=(i == 7))
            for i, (hd, d) in enumerate(zip(self.h_dilations, self.dilations)))
        self.layers.apply(add_weight_norms)
        self.end = nn.Conv2d(skip_channels, 2, 1, bias=bias)
        if zero_init:
            self.end.weight.data.zero_()
            if bias:
                self.end.bias.data.zero_()
        self._pack_cache = {}

    def _config(self, prec: str, height: int = 0) -> L.WnConfig:
        if height < 2:
            raise RuntimeError(f"WN2D: image height {height} unsupported (the 2-D pipeline needs >= 2 lines)")
        hd = (C.c_int * L.MAX_DEPTH)(*self.h_dilations)
        return L.WnConfig(1, self.aux_chs, self.dil_chs, self.res_chs, self.skp_chs, 8, 3, int(self.has_bias),
                          L.PREC_NAMES[prec], height, hd)

    def _tc_supported(self) -> bool:
        return bool(L.load().cmwg_wn_tc_supported(C.byref(self._config("fp32", 2))))

    def _prepare(self, prec: str, device, height: int = 0):
        # the packed weights do not depend on the height; the config does
        cfg, packed, ps = super()._prepare(prec, device, max(height, 2))
        return self._config(prec, height), packed, ps

    # ---- fused entry points -------------------------------------------------------------------------------
    def _fwd_image(self, img: Tensor, lines: int, y: Tensor, save: bool, prec: Optional[str] = None):
        """img: (B, Himg, W) contiguous fp32 whose first `lines` lines are the WN input (one channel).
        Returns (lst (B, 2, lines*W) = [log_s ; t], state)."""
        L.require_cuda(img, y, op="WN2D.forward")
        if prec is None:
            prec = precision.resolve(self._tc_supported(), training=save)
        cfg, packed, ps = self._prepare(prec, img.device, lines)
        lib = L.load()
        B, Himg, W = img.shape
        if y.shape[0] != B or y.shape[1] != self.aux_chs or y.shape[2] != W:
            raise RuntimeError(f"WN2D: conditioning shape {tuple(y.shape)} does not match input {(B, self.aux_chs, W)}")
        ycl = _cond_cache.get(y.float(), cfg)
        ws = torch.empty(int(lib.cmwg_wn_workspace_bytes(C.byref(cfg), B, W)), device=img.device, dtype=torch.uint8)
        saved = torch.empty(int(lib.cmwg_wn_saved_bytes(C.byref(cfg), B, W)), device=img.device,
                            dtype=torch.uint8) if save else None
        lst = torch.empty((B, 2, lines * W), device=img.device, dtype=torch.float32)
        L.check(lib.cmwg_wn_forward(C.byref(cfg), packed.data_ptr(), img.data_ptr(), Himg * W, ycl.data_ptr(), B, W,
                                    ws.data_ptr(), L.ptr(saved), lst.data_ptr(), L.stream_ptr(img.device)),
                "wn_forward(2d)")
        st = _WNState()
        st.cfg, st.packed, st.params, st.ycl, st.saved, st.B, st.T, st.prec = cfg, packed, ps, ycl, saved, B, W, prec
        return lst, st

    def _bwd_image(self, st: _WNState, img: Tensor, dlst: Tensor, dimg: Tensor, need_dy: bool):
        """Accumulates the WN input gradient into the first lines of dimg (B, Himg, W); returns (grads in
        self.parameters() order, dy or None)."""
        lib = L.load()
        dev = img.device
        B, Himg, W = img.shape
        grads, by_param = self._grads_struct()
        aux_p = lib.cmwg_wn_aux_padded(C.byref(st.cfg))
        dycl = torch.empty((B, W, aux_p), device=dev, dtype=torch.float32) if need_dy else None
        ws = torch.empty(int(lib.cmwg_wn_workspace_bytes(C.byref(st.cfg), B, W)), device=dev, dtype=torch.uint8)
        L.check(lib.cmwg_wn_backward(C.byref(st.cfg), C.byref(st.params), st.packed.data_ptr(), img.data_ptr(),
                                     Himg * W, st.ycl.data_ptr(), B, W, ws.data_ptr(), st.saved.data_ptr(),
                                     dlst.data_ptr(), dimg.data_ptr(), Himg * W, L.ptr(dycl), C.byref(grads),
                                     L.stream_ptr(dev)), "wn_backward(2d)")
        dy = None
        if need_dy:
            dy = torch.empty((B, self.aux_chs, W), device=dev, dtype=torch.float32)
            L.check(lib.cmwg_cond_unpack_grad(C.byref(st.cfg), dycl.data_ptr(), B, W, dy.data_ptr(), L.stream_ptr(dev)),
                    "cond_unpack_grad")
        return [by_param[id(p)] for p in self.parameters()], dy

    def _inverse_flow(self, z: Tensor, in_flip: bool, y: Tensor, prec: Optional[str] = None):
        """One flow of the synthesis direction: z (B, H, W) -> (x (B, H, W), lst (B, 2, (H-1)*W))."""
        L.require_cuda(z, y, op="WN2D.inverse")
        if prec is None:
            prec = precision.resolve(self._tc_supported(), training=False)
        B, H, W = z.shape
        cfg, packed, _ = self._prepare(prec, z.device, H - 1)
        lib = L.load()
        ycl = _cond_cache.get(y.float(), cfg)
        ws = torch.empty(int(lib.cmwg_wn_workspace_bytes(C.byref(cfg), B, W)), device=z.device, dtype=torch.uint8)
        state = torch.empty(int(lib.cmwg_wn_line_state_bytes(C.byref(cfg), B, W)), device=z.device, dtype=torch.uint8)
        lst = torch.empty((B, 2, (H - 1) * W), device=z.device, dtype=torch.float32)
        x = torch.empty_like(z)
        L.check(lib.cmwg_waveflow_inverse_flow(C.byref(cfg), packed.data_ptr(), z.data_ptr(), int(in_flip),
                                               ycl.data_ptr(), B, W, ws.data_ptr(), state.data_ptr(), lst.data_ptr(),
                                               x.data_ptr(), L.stream_ptr(z.device)), "waveflow_inverse_flow")
        return x, lst

    def forward(self, x, y):
        """x: (B, 1, H, W), y: (B, aux, W) -> (log_s, t) each (B, 1, H, W) (reference ``:128-135``)."""
        B, _, H, W = x.shape
        with grad_hint():
            lst = _WN2DFunction.apply(x.reshape(B, H, W), y, self, *self.parameters())
        return lst[:, 0].view(B, 1, H, W), lst[:, 1].view(B, 1, H, W)


class _WN2DFunction(torch.autograd.Function):
    """WN2D as an ordinary (activation-storing) autograd node, what the reference gets from autograd."""

    @staticmethod
    def forward(ctx, img, y, wn, *params):
        need = any(ctx.needs_input_grad) and graph_possible()
        img = img.detach().contiguous().float()
        lst, st = wn._fwd_image(img, img.shape[1], y.detach(), save=need)
        ctx.wn, ctx.st = wn, st
        ctx.save_for_backward(img)
        return lst

    @staticmethod
    def backward(ctx, dlst):
        (img,) = ctx.saved_tensors
        dimg = torch.zeros_like(img)
        grads, dy = ctx.wn._bwd_image(ctx.st, img, dlst.contiguous(), dimg, need_dy=ctx.needs_input_grad[1])
        ctx.st = None
        return (dimg if ctx.needs_input_grad[0] else None, dy, None) + tuple(grads)


def _affine(inp: Tensor, in_flip: bool, lst: Optional[Tensor], out: Tensor, out_flip: bool, inverse: bool,
            line_begin: int = 0, line_count: Optional[int] = None) -> Tensor:
    B, H, W = inp.shape
    L.check(L.load().cmwg_waveflow_affine(inp.data_ptr(), int(in_flip), L.ptr(lst), out.data_ptr(), int(out_flip), B, H,
                                          W, line_begin, H - line_begin if line_count is None else line_count,
                                          int(inverse), L.stream_ptr(inp.device)), "waveflow_affine")
    return out


class _WaveFlowStep(torch.autograd.Function):
    """One flow of ``WaveFlow.forward_computation`` (reference ``model/waveflow.py:203-211``) as a single node:
    (img, y) -> (new image, log_s.sum((1,2,3))).  ``flip``: the height flip of ``cat(xout.flip(2), x0)``; without it
    the output is ``cat(x0, xout)`` (the input of the optional 1x1 conv, :213)."""

    @staticmethod
    def forward(ctx, img, y, wn, flip, *params):
        need = any(ctx.needs_input_grad) and graph_possible()
        img = img.detach().contiguous().float()
        B, H, W = img.shape
        lst, st = wn._fwd_image(img, H - 1, y.detach(), save=need)
        out = _affine(img, False, lst, torch.empty_like(img), flip, False)
        logdet = ops.sum_per_batch(lst[:, :1])
        ctx.wn, ctx.st, ctx.flip = wn, st, flip
        ctx.save_for_backward(img, lst)
        return out, logdet

    @staticmethod
    def backward(ctx, dout, dlogdet):
        img, lst = ctx.saved_tensors
        B, H, W = img.shape
        dimg = torch.empty_like(img)
        dlst = torch.empty_like(lst)
        dout = dout.contiguous().float()
        dld = None if dlogdet is None else dlogdet.contiguous().float()
        L.check(L.load().cmwg_waveflow_affine_bwd(img.data_ptr(), lst.data_ptr(), dout.data_ptr(), int(ctx.flip),
                                                  L.ptr(dld), dimg.data_ptr(), dlst.data_ptr(), B, H, W,
                                                  L.stream_ptr(img.device)), "waveflow_affine_bwd")
        grads, dy = ctx.wn._bwd_image(ctx.st, img, dlst, dimg, need_dy=ctx.needs_input_grad[1])
        ctx.st = None
        return (dimg if ctx.needs_input_grad[0] else None, dy, None, None) + tuple(grads)


class _DenseUpsampleFunction(torch.autograd.Function):
    """ReplicationPad1d((0, rpad)) -> weight-normed dense ConvTranspose1d -> LeakyReLU (``model/waveflow.py:169-175``)."""

    @staticmethod
    def forward(ctx, h, g, v, bias, stride, pad, rpad, slope):
        L.require_cuda(h, v, op="WaveFlow.upsampler")
        lib = L.load()
        h = h.detach().contiguous().float()
        B, Cc, Fr = h.shape
        if v.shape[0] != Cc or v.shape[1] != Cc:
            raise RuntimeError("WaveFlow upsampler: expected a (n_mels, n_mels, K) transposed-conv weight")
        K = v.shape[-1]
        Tout = (Fr + rpad - 1) * stride - 2 * pad + K
        y = torch.empty((B, Cc, Tout), device=h.device, dtype=torch.float32)
        ws = torch.empty(int(lib.cmwg_upsample_dense_workspace(Cc, K)), device=h.device, dtype=torch.uint8)
        L.check(lib.cmwg_upsample_dense_fwd(h.data_ptr(), L.ptr(g), v.data_ptr(), L.ptr(bias), B, Cc, Fr, K, stride, pad,
                                            rpad, slope, y.data_ptr(), ws.data_ptr(), L.stream_ptr(h.device)),
                "upsample_dense_fwd")
        ctx.save_for_backward(h, g, v, y)
        ctx.cfg = (stride, pad, rpad, slope, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        h, g, v, y = ctx.saved_tensors
        stride, pad, rpad, slope, has_bias = ctx.cfg
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("WaveFlow upsampler: gradient w.r.t. the mel input is not implemented")
        lib = L.load()
        B, Cc, Fr = h.shape
        K = v.shape[-1]
        dg = torch.empty_like(g) if g is not None else None
        dv = torch.empty_like(v)
        db = torch.empty((Cc,), device=h.device, dtype=torch.float32) if has_bias else None
        ws = torch.empty(int(lib.cmwg_upsample_dense_workspace(Cc, K)), device=h.device, dtype=torch.uint8)
        L.check(lib.cmwg_upsample_dense_bwd(h.data_ptr(), L.ptr(g), v.data_ptr(), y.data_ptr(),
                                            dy.contiguous().float().data_ptr(), B, Cc, Fr, K, stride, pad, rpad, slope,
                                            L.ptr(dg), dv.data_ptr(), L.ptr(db), ws.data_ptr(), L.stream_ptr(h.device)),
                "upsample_dense_bwd")
        return None, dg, dv, db, None, None, None, None


class WaveFlow(FlowBase):
    """Reference ``model/waveflow.py:154-265``: squeeze to a (n_group, T/n_group) image, ``flows`` x (2-D WN on rows
    0..H-2 -> affine transform of rows 1..H-1 -> height flip or invertible 1x1 conv over the rows)."""

    def __init__(self, flows, n_group, n_mels, use_conv1x1, memory_efficient, reverse_mode=False, **kwargs):
        super().__init__(256, reverse_mode)
        self.flows = flows
        self.n_group = n_group
        self.n_mels = n_mels
        self.sub_sr = self._hop_length // n_group

        self.upsampler = nn.Sequential(
            nn.ReplicationPad1d((0, 1)),
            nn.ConvTranspose1d(n_mels, n_mels, self.sub_sr * 2 + 1, self.sub_sr, padding=self.sub_sr // 2),
            nn.LeakyReLU(0.4, True))
        self.upsampler.apply(add_weight_norms)

        self.WNs = nn.ModuleList()
        if use_conv1x1:
            self.invconv1x1 = nn.ModuleList()
        for _ in range(flows):
            self.WNs.append(WN2D(n_group, n_mels, **kwargs))
            if use_conv1x1:
                self.invconv1x1.append(InvertibleConv1x1(n_group, memory_efficient=memory_efficient,
                                                         reverse_mode=reverse_mode))
        self._graphs = {}  # synthesis-direction CUDA graphs, keyed by shapes + weight versions

    def _upsample_h(self, h):
        up = self.upsampler[1]
        g, v, b = WN._gvb(up)
        return _DenseUpsampleFunction.apply(h.float(), g, v, b, up.stride[0], up.padding[0],
                                            self.upsampler[0].padding[1], self.upsampler[2].negative_slope)

    def forward_computation(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        L.require_cuda(x, h, op="WaveFlow.forward")
        y = self._upsample_h(h)
        batch = x.size(0)
        img = _SqueezeFunction.apply(x, self.n_group, False)  # (B, n_group, T/n_group): image[h, w] = x[w*n_group + h]
        y = y[..., :img.size(-1)]
        convs = self.invconv1x1 if hasattr(self, "invconv1x1") else [None] * self.flows
        logdet = None
        for wn, invconv in zip(self.WNs, convs):
            with grad_hint():
                img, term = _WaveFlowStep.apply(img, y, wn, invconv is None, *wn.parameters())
            if invconv is not None:
                img, log_det_w = invconv(img)
                term = term + log_det_w
            logdet = term if logdet is None else logdet + term
        z = _SqueezeFunction.apply(img, self.n_group, True)
        return z.view(batch, -1), logdet

    def _reverse_eager(self, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        y = self._upsample_h(h)
        batch = z.size(0)
        img = ops.squeeze(z, self.n_group, False)
        y = y[..., :img.size(-1)]
        convs = self.invconv1x1 if hasattr(self, "invconv1x1") else [None] * self.flows
        logdet = None
        for wn, invconv in zip(self.WNs[::-1], convs[::-1]):
            if invconv is not None:
                img, log_det_w = invconv.reverse(img)
                logdet = log_det_w.repeat(batch) if logdet is None else logdet + log_det_w
            img, lst = wn._inverse_flow(img, invconv is None, y)
            term = ops.sum_per_batch(lst[:, :1], scale=-1.0)
            logdet = term if logdet is None else logdet + term
        x = ops.squeeze(img, self.n_group, True)
        return x.view(batch, -1), logdet

    def _graph_key(self, z: Tensor, h: Tensor):
        return (tuple(z.shape), tuple(h.shape), z.device.index, precision.get_precision(),
                torch.backends.cudnn.allow_tf32, pack_generation(),
                tuple((p.data_ptr(), p._version) for p in self.parameters()))

    def reverse_computation(self, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        """Synthesis direction (reference ``:221-261``).  Runs without building an autograd graph: the row-recurrent
        loop is one fused device-side sequence per flow (~1200 small launches per flow, launch bound), so from the
        second call with the same shapes and weights on the whole direction is replayed as ONE CUDA graph
        (``CMWG_GRAPHS=0`` disables that)."""
        L.require_cuda(z, h, op="WaveFlow.reverse")
        if torch.is_grad_enabled() and (z.requires_grad or h.requires_grad):
            raise NotImplementedError("WaveFlow.reverse_computation is inference-only (no autograd through the row loop)")
        with torch.no_grad():
            z = z.detach().float().contiguous()
            h = h.detach().float().contiguous()
            if os.environ.get("CMWG_GRAPHS", "1") == "0" or torch.cuda.is_current_stream_capturing():
                return self._reverse_eager(z, h)
            key = self._graph_key(z, h)
            if key not in self._graphs:
                # first sight of this (shape, weights) combination: plain launches (this also warms the tensor-map
                # and weight-pack caches); a repeat is worth a capture
                self._graphs = {k: v for k, v in self._graphs.items() if v is not None}  # drop stale "seen" marks
                while len(self._graphs) >= 2:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = None
                return self._reverse_eager(z, h)
            if self._graphs[key] is None:
                sz, sh = z.clone(), h.clone()
                torch.cuda.synchronize(z.device)
                graph = torch.cuda.CUDAGraph()
                _cond_cache.clear()  # conditioning slabs cached from eager calls must be re-packed INSIDE the graph
                with torch.cuda.graph(graph):
                    ox, old = self._reverse_eager(sz, sh)
                _cond_cache.clear()  # ... and slabs living in the graph's private pool must not serve eager calls
                # the capture baked in the addresses of every WN's packed-weight buffer (allocated by earlier eager calls,
                # outside the graph's pool): hold them for as long as the graph lives
                pins = [ent[1] for wn in self.WNs for ent in wn._pack_cache.values()]
                self._graphs[key] = (graph, sz, sh, ox, old, pins)
            graph, sz, sh, ox, old, _pins = self._graphs[key]
            sz.copy_(z)
            sh.copy_(h)
            graph.replay()
            return ox.clone(), old.clone()
