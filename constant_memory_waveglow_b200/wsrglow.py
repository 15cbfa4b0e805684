"""WSRGlow (audio super-resolution flow): host-side mirror of the reference's ``model/wsrglow.py``.

``WSRGlow`` (:21-56) is a ``WaveGlow`` with 12 flows, ``n_group = hop = 8 * upsample_rate``, early outputs
4 / 2 and a 3659-channel conditioning built from the low-rate signal by ``_get_cond`` (:37-50): mu-law code
embedding (3200 rows), 9 STFT magnitudes and 9 x 50 phase-embedding rows, at one conditioning column per
squeezed time step (upsample factor 1).  The flow itself -- 1x1 convs, couplings, the WN stack whose ``V``
conv is now 3659 -> 2*Cd*depth and dominates the FLOPs (78 %) -- runs on the same libcmwg_b200.so kernels as
WaveGlow; the conditioning front end is a handful of index / elementwise torch ops (no cuFFT: the 16-point
STFT is written out as a windowed DFT so that it stays a plain elementwise + reduction computation).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .waveglow import WaveGlow


class AngleEmbedding(nn.Module):
    """Reference ``model/wsrglow.py:8-18``: quantise an angle in [-pi, pi] to ``embed_num`` codes."""

    def __init__(self, embed_num, hidden_dim):
        super().__init__()
        self.embed_num = embed_num
        self.embed = nn.Embedding(num_embeddings=embed_num, embedding_dim=hidden_dim)

    def forward(self, index):
        embed_num = self.embed_num
        index = ((index / torch.pi + 1) * 0.5 * (embed_num - 1)).long()
        return self.embed(index)


class MuLawEncoding(nn.Module):
    """torchaudio.transforms.MuLawEncoding(quantization_channels) without the torchaudio dependency."""

    def __init__(self, quantization_channels: int = 256):
        super().__init__()
        self.quantization_channels = quantization_channels

    def forward(self, x):
        mu = torch.tensor(self.quantization_channels - 1.0, dtype=x.dtype, device=x.device)
        x_mu = torch.sign(x) * torch.log1p(mu * torch.abs(x)) / torch.log1p(mu)
        return ((x_mu + 1) / 2 * mu + 0.5).to(torch.int64)


class WSRGlow(WaveGlow):
    """Reference ``model/wsrglow.py:21-56`` (same constructor, attributes and state-dict keys:
    ``mu_enc.1.weight``, ``angle_embed.embed.weight``, buffer ``window``)."""

    def __init__(self, upsample_rate: int = 2, memory_efficient: bool = False, **kwargs) -> None:
        super().__init__(12, 8 * upsample_rate, 4, 2, 8 * upsample_rate, 8 * 400 + 51 * 9,
                         memory_efficient=memory_efficient, **kwargs)
        self.mu_enc = nn.Sequential(MuLawEncoding(256), nn.Embedding(256, 400))
        self.angle_embed = AngleEmbedding(embed_num=120, hidden_dim=50)
        self.n_fft = 16
        self.hop_length = 8
        self.register_buffer('window', torch.hann_window(self.n_fft))
        k = torch.arange(self.n_fft // 2 + 1, dtype=torch.float64).view(-1, 1)
        n = torch.arange(self.n_fft, dtype=torch.float64).view(1, -1)
        ang = 2 * math.pi * k * n / self.n_fft
        sin = torch.sin(ang)
        sin[0] = 0.0                   # DC and Nyquist bins of a real signal are exactly real: a real FFT returns
        sin[self.n_fft // 2] = 0.0     # imag = +0.0 there, and atan2(+0, re < 0) = +pi picks the LAST phase code
        self.register_buffer('_dft_cos', torch.cos(ang).float(), persistent=False)
        self.register_buffer('_dft_sin', sin.float(), persistent=False)

    def _stft(self, c):
        """torch.stft(pad_reflect(c, 4), n_fft=16, hop=8, window, center=False) as a windowed DFT:
        returns (re, im), each (B, 9, frames)."""
        xp = F.pad(c.unsqueeze(1), (4, 4), mode='reflect').squeeze(1)
        frames = xp.unfold(-1, self.n_fft, self.hop_length) * self.window          # (B, frames, 16)
        fr = frames.unsqueeze(1)                                                    # (B, 1, frames, 16)
        re = (fr * self._dft_cos.view(1, -1, 1, self.n_fft)).sum(-1)
        s = (fr * self._dft_sin.view(1, -1, 1, self.n_fft)).sum(-1)
        im = torch.zeros_like(s) - s   # (+0) - (+-0) = +0: keeps the sign convention of the real FFT
        return re, im

    def _get_cond(self, c):
        c = c.clip_(-1, 1)
        c_emb = self.mu_enc(c).view(c.shape[0], -1, 8 * 400).transpose(1, 2)
        re, im = self._stft(c)
        mag = torch.sqrt(re * re + im * im)
        phase_emb = self.angle_embed(torch.atan2(im, re)).permute(0, 1, 3, 2).reshape(re.shape[0], 50 * 9, -1)
        return torch.cat([c_emb, mag, phase_emb], dim=1)

    def forward_computation(self, x, h):
        return super().forward_computation(x, self._get_cond(h))

    def reverse_computation(self, z, h):
        return super().reverse_computation(z, self._get_cond(h))
