"""MRWaveGlow -- the multi-resolution wiring of the same flow primitives (reference
``model/mr_waveglow.py:14-131``).

The squeezed signal (B, n_group, T') is split ``levels - 1`` times into a detail band ``odd - even`` and a coarse band
``(even + odd) / 2`` over the channel axis (a Haar pair).  Every detail band goes through ``flows`` steps of invertible
1x1 conv + affine coupling whose WN is conditioned on [coarse band ; upsampled mel] (or the coarse band alone for
super-resolution); the last coarse band goes through ``prior_flows`` steps conditioned on the mel.  Only the wiring is new:
the 1x1 convs, the couplings, their WNs and the constant-memory backward are the kernels of ``efficient_modules.py`` /
``waveglow.py`` (the level couplings receive a conditioning that depends on the signal, so their backward also returns
``dy``, ``cmwg_wn_backward``'s conditioning gradient).  The mel is upsampled by linear interpolation
(``F.interpolate(mode='linear')``, parameter-free: plain torch), as in the reference (``:133-134``).

Kept from the reference: constructor signature and defaults, attribute / state-dict names (``prior_invconv1x1``,
``prior_WNs``, ``invconv1x1_list.{level}.{k}``, ``WNs_list.{level}.{k}``) and its quirk that the level 1x1 convs are ALWAYS
memory-efficient (``InvertibleConv1x1(in_channels, in_channels)`` passes the channel count as the flag, ``:46``).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .base import FlowBase
from .efficient_modules import AffineCouplingBlock, InvertibleConv1x1
from .waveglow import WN

__all__ = ["MRWaveGlow"]


class MRWaveGlow(FlowBase):
    def __init__(self, prior_flows, n_group, hop_size, n_mels, memory_efficient, levels=3, flows=4,
                 super_resolution=False, reverse_mode=False, **kwargs):
        super().__init__(hop_size, reverse_mode)
        self.flows, self.prior_flows, self.levels = flows, prior_flows, levels
        self.n_group, self.n_mels, self.super_resolution = n_group, n_mels, super_resolution
        self.upsample_factor = hop_size // n_group

        self.prior_invconv1x1 = nn.ModuleList()
        self.prior_WNs = nn.ModuleList()
        self.invconv1x1_list = nn.ModuleList()
        self.WNs_list = nn.ModuleList()

        band = n_group
        for _ in range(levels - 1):
            band //= 2                       # channels of this level's detail band (and of its coarse band)
            aux = band + (0 if super_resolution else n_mels)
            self.invconv1x1_list.append(nn.ModuleList(InvertibleConv1x1(band, True) for _ in range(flows)))
            self.WNs_list.append(nn.ModuleList(
                AffineCouplingBlock(WN, memory_efficient=memory_efficient, reverse_mode=reverse_mode,
                                    in_channels=band // 2, aux_channels=aux, **kwargs) for _ in range(flows)))
        for _ in range(prior_flows):
            self.prior_invconv1x1.append(InvertibleConv1x1(band, memory_efficient=memory_efficient,
                                                           reverse_mode=reverse_mode))
            self.prior_WNs.append(AffineCouplingBlock(WN, memory_efficient=memory_efficient, in_channels=band // 2,
                                                      aux_channels=n_mels, reverse_mode=reverse_mode, **kwargs))

    # ---- pieces -------------------------------------------------------------------------------------------------
    def _upsample_h(self, h: Tensor) -> Tensor:
        return F.interpolate(h, scale_factor=self.upsample_factor, mode="linear")

    def _squeezed(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        y = self._upsample_h(h)
        x = x.view(x.size(0), -1, self.n_group).transpose(1, 2)
        assert x.size(2) <= y.size(2)
        return x, y[..., :x.size(2)]

    def _level_cond(self, coarse: Tensor, y: Tensor) -> Tensor:
        return coarse if self.super_resolution else torch.cat((coarse, y), 1)

    @staticmethod
    def _steps(convs, couplings, x: Tensor, cond: Tensor, inverse: bool):
        """`flows` steps over one band in the given direction; returns (band, summed log-determinant)."""
        total = 0
        order = range(len(convs) - 1, -1, -1) if inverse else range(len(convs))
        for k in order:
            if inverse:
                x, log_s = couplings[k].reverse(x, cond)
                x, ld_w = convs[k].reverse(x)
            else:
                x, ld_w = convs[k](x)
                x, log_s = couplings[k](x, cond)
            total = total + ld_w + log_s.sum((1, 2))
        return x, total

    # ---- directions ---------------------------------------------------------------------------------------------
    def forward_computation(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        B = x.size(0)
        coarse, y = self._squeezed(x, h)
        bands: List[Tensor] = []
        logdet = 0
        for level in range(self.levels - 1):
            even, odd = coarse[:, ::2], coarse[:, 1::2]
            detail, coarse = odd - even, (even + odd) * 0.5
            detail, ld = self._steps(self.invconv1x1_list[level], self.WNs_list[level], detail,
                                     self._level_cond(coarse, y), inverse=False)
            logdet = logdet + ld
            bands.append(detail)
        coarse, ld = self._steps(self.prior_invconv1x1, self.prior_WNs, coarse.contiguous(), y, inverse=False)
        bands.append(coarse)
        return torch.cat(bands, 1).transpose(1, 2).contiguous().view(B, -1), logdet + ld

    def reverse_computation(self, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        B = z.size(0)
        z, y = self._squeezed(z, h)
        details: List[Tensor] = []
        for _ in range(self.levels - 1):
            d, z = z.chunk(2, 1)
            details.append(d.clone())            # the memory-efficient steps consume their input
        coarse, logdet = self._steps(self.prior_invconv1x1, self.prior_WNs, z.contiguous(), y, inverse=True)
        for level in range(self.levels - 2, -1, -1):
            detail, ld = self._steps(self.invconv1x1_list[level], self.WNs_list[level], details.pop(),
                                     self._level_cond(coarse, y), inverse=True)
            logdet = logdet + ld
            even, odd = coarse - detail * 0.5, coarse + detail * 0.5
            coarse = torch.stack((even, odd), 2).view(B, -1, even.size(2))
        return coarse.transpose(1, 2).contiguous().view(B, -1), logdet
