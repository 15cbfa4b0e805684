"""MelGlow -- WaveGlow's flow with a location-variable-convolution transform (reference
``model/melglow.py:13-258``).

What is native here and what is not: the flow itself -- invertible 1x1 convs, affine couplings, their constant-memory
backward, early outputs -- runs on this package's CUDA kernels; the TRANSFORM ``WN_LVC`` is a ``transform_type`` other than
this package's ``WN``, so ``AffineCouplingBlock`` calls it as a module and differentiates it with autograd (the generic-F
path of ``efficient_modules.py``, as the reference does for every transform).  ``WN_LVC`` below is written in plain PyTorch
ops: a kernel predictor at frame rate (grouped 1x1 convs + BatchNorm + tanh) emits one (2*Cd, Cr, radix) kernel per frame and
layer, and each layer applies its frame's kernel to that frame's ``hop/n_group`` columns.  Its channel counts are small
(48 in ``configs/melglow_LJ_speech.json``) and it is not one of BASELINE's configs; sm_100a kernels for it are future work.

The location-variable convolution is computed tap by tap -- ``z[b,:,s,:] += W[b,s,:,:,k] @ x_pad[b,:,frame s shifted by k*dil]`` --
instead of the reference's unfold + grouped ``conv1d`` (``:80-87``); the sums are the same up to fp32 ordering.  Constructor
signatures, attribute names and state-dict keys are the reference's.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _lib as L
from .base import FlowBase
from .efficient_modules import AffineCouplingBlock, InvertibleConv1x1
from .utils import add_weight_norms
from .waveglow import fused_gate

__all__ = ["MelGlow", "WN_LVC", "NonCausalLayerLVC", "Predictor"]


class Predictor(nn.Module):
    """Frame-rate network that predicts the layers' kernels (reference ``model/melglow.py:13-49``): ``groups`` = one group
    per WN layer, so every layer's kernels come from its own slice of the hidden state."""

    def __init__(self, in_channels, out_channels, hidden_channels, layers, bias, groups):
        super().__init__()
        self.groups = groups
        width = hidden_channels * groups

        def unit():
            return [nn.Conv1d(width, width, 1, bias=bias, groups=groups), nn.BatchNorm1d(width), nn.Tanh()]

        self.start = nn.Sequential(nn.Conv1d(in_channels, width, 1, bias=bias), nn.BatchNorm1d(width), nn.Tanh())
        self.end = nn.Conv1d(width, out_channels * groups, 1, bias=bias, groups=groups)
        self.res_blocks = nn.ModuleList(nn.Sequential(*unit(), *unit()) for _ in range(layers))

    def forward(self, y: Tensor) -> Tensor:
        s = self.start(y)
        for block in self.res_blocks:
            s = s + block(s)
        return self.end(s)


class _LVCGate(torch.autograd.Function):
    """Location-variable dilated conv + gate as one kernel each way (``cmwg_lvc_gate_forward`` / ``_backward``,
    csrc/lvc.cu): reference ``model/melglow.py:72-85`` (pad, unfold, grouped ``F.conv1d`` with one kernel per frame,
    ``fused_gate``).  x (B, Cr, T), weights (B, frames, 2Cd, Cr, radix) -> g (B, Cd, T)."""

    @staticmethod
    def forward(ctx, x, weights, dilation):
        L.require_cuda(x, weights, op="WN_LVC")
        x = x.detach().float().contiguous()
        w = weights.detach().float().contiguous()
        B, frames, o2, cr, radix = w.shape
        T = x.shape[2]
        g = torch.empty((B, o2 // 2, T), device=x.device, dtype=torch.float32)
        L.check(L.load().cmwg_lvc_gate_forward(x.data_ptr(), w.data_ptr(), B, T, frames, o2 // 2, cr, radix, int(dilation),
                                               g.data_ptr(), L.stream_ptr(x.device)), "lvc_gate_forward")
        ctx.save_for_backward(x, w)
        ctx.dilation = int(dilation)
        return g

    @staticmethod
    def backward(ctx, dg):
        x, w = ctx.saved_tensors
        B, frames, o2, cr, radix = w.shape
        T = x.shape[2]
        dg = dg.float().contiguous()
        dz = torch.empty((B, o2, T), device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        L.check(L.load().cmwg_lvc_gate_backward(x.data_ptr(), w.data_ptr(), dg.data_ptr(), B, T, frames, o2 // 2, cr, radix,
                                                ctx.dilation, dz.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                                L.stream_ptr(x.device)), "lvc_gate_backward")
        return dx, dw, None


class NonCausalLayerLVC(nn.Module):
    """One layer (reference ``model/melglow.py:52-92``): location-variable dilated conv -> gate -> ``W_o`` 1x1 -> residual / skip."""

    def __init__(self, dilation, dilation_channels, residual_channels, skip_channels, radix, bias, last_layer=False):
        super().__init__()
        self.dilation = dilation
        self.padding = dilation * (radix - 1) // 2
        self.chs_split = [skip_channels] if last_layer else [residual_channels, skip_channels]
        self.W_o = nn.Conv1d(dilation_channels, sum(self.chs_split), 1, bias=bias)

    def _lvc_gate_torch(self, x: Tensor, weights: Tensor) -> Tensor:
        """The same computation as torch ops (tap-by-tap einsum): what tests/test_melglow.py checks the kernels against;
        not on the product path."""
        B, frames, cout, cin, radix = weights.shape
        T = x.shape[2]
        span = T // frames
        xp = F.pad(x, (self.padding, self.padding))
        z = None
        for k in range(radix):
            tap = xp[:, :, k * self.dilation:k * self.dilation + T].reshape(B, cin, frames, span)
            term = torch.einsum("bsoc,bcst->bost", weights[..., k], tap)
            z = term if z is None else z + term
        zw, zv = z.reshape(B, cout, T).chunk(2, 1)
        return fused_gate(zw, zv)

    def forward(self, x: Tensor, weights: Tensor):
        """x (B, Cr, T); weights (B, frames, 2*Cd, Cr, radix), frame s owns columns [s*T/frames, (s+1)*T/frames)."""
        out = self.W_o(_LVCGate.apply(x, weights, self.dilation))     # raises for CPU tensors: there is no CPU path
        if len(self.chs_split) == 1:
            return None, out
        res, skip = out.split(self.chs_split, 1)
        return res + x, skip


class WN_LVC(nn.Module):
    """Reference ``model/melglow.py:95-159``.  The location-variable convolutions and gates run in csrc/lvc.cu; the kernel
    predictor and the 1x1 ``W_o`` / ``start`` / ``end`` convolutions are PyTorch modules."""

    def __init__(self, in_channels, aux_channels, depth, dilation_channels, residual_channels, skip_channels,
                 predict_channels, predict_layers, radix, bias, zero_init=True):
        super().__init__()
        self.dilations = [2 ** i for i in range(depth)]
        self.in_chs, self.res_chs, self.dil_chs, self.skp_chs, self.rdx = \
            in_channels, residual_channels, dilation_channels, skip_channels, radix
        self.r_field = sum(self.dilations) + 1

        self.start = nn.Conv1d(in_channels, residual_channels, 1, bias=bias)
        self.start.apply(add_weight_norms)
        self.layers = nn.ModuleList(
            NonCausalLayerLVC(d, dilation_channels, residual_channels, skip_channels, radix, bias,
                              last_layer=(i == depth - 1)) for i, d in enumerate(self.dilations))
        self.layers.apply(add_weight_norms)
        self.end = nn.Conv1d(skip_channels, in_channels * 2, 1, bias=bias)
        if zero_init:
            self.end.weight.data.zero_()
            if bias:
                self.end.bias.data.zero_()
        self.pred = Predictor(aux_channels, 2 * dilation_channels * residual_channels * radix, predict_channels,
                              predict_layers, bias, depth)

    def forward(self, x: Tensor, y: Tensor):
        h = self.start(x)
        B, frames = y.shape[0], y.shape[2]
        depth = len(self.dilations)
        kernels = self.pred(y).view(B, depth, -1, frames).permute(1, 0, 3, 2)   # (depth, B, frames, 2*Cd*Cr*radix)
        total = 0
        for layer, w in zip(self.layers, kernels):
            h, skip = layer(h, w.reshape(B, frames, 2 * self.dil_chs, self.res_chs, self.rdx))
            total = total + skip
        return self.end(total).chunk(2, 1)


class MelGlow(FlowBase):
    """Reference ``model/melglow.py:162-258``: the mel reaches the transforms at FRAME rate (no upsampler)."""

    def __init__(self, flows, n_group, n_early_every, n_early_size, hop_size, n_mels, memory_efficient,
                 reverse_mode=False, **kwargs):
        super().__init__(hop_size, reverse_mode=reverse_mode)
        self.flows, self.n_group, self.n_mels = flows, n_group, n_mels
        self.n_early_every, self.n_early_size = n_early_every, n_early_size
        self.mem_efficient = memory_efficient
        self.upsample_factor = self._hop_length // n_group
        self.invconv1x1 = nn.ModuleList()
        self.WNs = nn.ModuleList()
        self.z_split_sizes: List[int] = []
        remaining = n_group
        for k in range(flows):
            if k and k % n_early_every == 0:
                remaining -= n_early_size
                self.z_split_sizes.append(n_early_size)
            self.invconv1x1.append(InvertibleConv1x1(remaining, memory_efficient=memory_efficient,
                                                     reverse_mode=reverse_mode))
            self.WNs.append(AffineCouplingBlock(WN_LVC, memory_efficient=memory_efficient, reverse_mode=reverse_mode,
                                                in_channels=remaining // 2, aux_channels=n_mels, **kwargs))
        self.z_split_sizes.append(remaining)

    def _squeezed(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        hop = self._hop_length
        x = x[:, :x.shape[1] // hop * hop]
        x = x.view(x.size(0), -1, self.n_group).transpose(1, 2)
        return x, h[..., :x.shape[2] // self.upsample_factor]

    def forward_computation(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        B = x.size(0)
        x, y = self._squeezed(x, h)
        early: List[Tensor] = []
        logdet = 0
        for k in range(self.flows):
            if k and k % self.n_early_every == 0:
                early.append(x[:, :self.n_early_size])
                x = x[:, self.n_early_size:]
                if self.mem_efficient:
                    x = x.clone()                 # the memory-efficient steps consume their input
            x, ld_w = self.invconv1x1[k](x)
            x, log_s = self.WNs[k](x, y)
            logdet = logdet + ld_w + log_s.sum((1, 2))
        assert x.shape[1] == self.z_split_sizes[-1]
        early.append(x)
        return torch.cat([e.transpose(1, 2) for e in early], 2).reshape(B, -1), logdet

    def reverse_computation(self, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        B = z.size(0)
        z, y = self._squeezed(z, h)
        parts = [p.clone() if self.mem_efficient else p for p in z.split(self.z_split_sizes, 1)]
        z = parts.pop()
        logdet = 0
        for k in range(self.flows - 1, -1, -1):
            z, log_s = self.WNs[k].reverse(z, y)
            z, ld_w = self.invconv1x1[k].reverse(z)
            logdet = logdet + ld_w + log_s.sum((1, 2))
            if k and k % self.n_early_every == 0:
                z = torch.cat((parts.pop(), z), 1)
        return z.transpose(1, 2).contiguous().view(B, -1), logdet
