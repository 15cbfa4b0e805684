"""Conditioners: the step right before the flow in training and inference (reference ``model/condition.py``).

``MelSpec(sr, n_fft, hop_length, **kwargs)`` mirrors ``model/condition.py:7-19`` -- reflection padding by
``(n_fft/2 - hop/2, n_fft/2 + hop/2)``, torchaudio's ``MelSpectrogram(center=False)`` (periodic Hann window, power 2,
HTK mel scale, no normalisation unless asked for) and ``log(mel + 1e-7)`` -- as ONE fused CUDA kernel
(``cmwg_melspec_fwd``, csrc/melspec.cu): the padded signal, the complex STFT and the power spectrogram never reach
HBM.  torchaudio is not imported; the filterbank is rebuilt from its published definition
(``torchaudio.functional.melscale_fbanks``) and the buffers keep the reference's state-dict keys
(``mel.1.spectrogram.window``, ``mel.1.mel_scale.fb``, persistent buffers as in torchaudio, so a Lightning
checkpoint's ``conditioner.mel.1.*`` entries load here and vice versa).
"""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn

from . import _lib as L

__all__ = ["MelSpec", "mel_filterbank"]


def _hz_to_mel(f: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + f / 700.0)
    # slaney: linear below 1 kHz, logarithmic above
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f / f_sp


def _mel_to_hz(m: Tensor, scale: str) -> Tensor:
    if scale == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    hz = f_sp * m
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    above = m >= min_log_mel
    hz[above] = 1000.0 * torch.exp(logstep * (m[above] - min_log_mel))
    return hz


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int, norm=None,
                   mel_scale: str = "htk") -> Tensor:
    """(n_freqs, n_mels) triangular filters, fp32 -- the definition of torchaudio.functional.melscale_fbanks:
    bin centres linspace(0, sr // 2, n_freqs); n_mels + 2 band edges equally spaced on the mel axis between
    f_min and f_max; band m rises from edge m to edge m+1 and falls to edge m+2."""
    if norm not in (None, "slaney"):
        raise ValueError('norm must be one of None or "slaney"')
    if mel_scale not in ("htk", "slaney"):
        raise ValueError('mel_scale should be one of "htk" or "slaney".')
    bins = torch.linspace(0, sample_rate // 2, n_freqs)
    edges_mel = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    edges = _mel_to_hz(edges_mel, mel_scale)
    width = edges[1:] - edges[:-1]                              # (n_mels + 1)
    dist = edges.unsqueeze(0) - bins.unsqueeze(1)               # (n_freqs, n_mels + 2)
    falling = -dist[:, :-2] / width[:-1]
    rising = dist[:, 2:] / width[1:]
    fb = torch.clamp(torch.minimum(falling, rising), min=0.0)
    if norm == "slaney":
        fb = fb * (2.0 / (edges[2:n_mels + 2] - edges[:n_mels])).unsqueeze(0)
    return fb


class _Buffers(nn.Module):
    """Holds one buffer under the attribute name torchaudio uses."""

    def __init__(self, name: str, value: Tensor):
        super().__init__()
        self.register_buffer(name, value)


class _MelSpectrogramState(nn.Module):
    """Parameter container laid out like torchaudio.transforms.MelSpectrogram (``spectrogram.window``,
    ``mel_scale.fb``) so module paths and state-dict keys match the reference's ``MelSpec.mel[1]``."""

    def __init__(self, window: Tensor, fb: Tensor):
        super().__init__()
        self.spectrogram = _Buffers("window", window)
        self.mel_scale = _Buffers("fb", fb)


class _ReflectionPadSpec(nn.Module):
    """Records the padding of ``nn.ReflectionPad1d`` (slot 0 of the reference's Sequential); the padding itself is
    index arithmetic inside the kernel."""

    def __init__(self, padding):
        super().__init__()
        self.padding = tuple(padding)

    def extra_repr(self) -> str:
        return f"{self.padding}"


class MelSpec(nn.Module):
    """Log-mel conditioner, x (B, T) or (T,) -> (B, n_mels, T // hop + 1)  (reference ``model/condition.py:7-19``).

    Keyword arguments are those of ``torchaudio.transforms.MelSpectrogram`` that the reference's configs use or
    that have a direct meaning here: ``f_min``, ``f_max``, ``n_mels``, ``win_length`` (<= n_fft, centred zero
    padding), ``power`` (1 or 2), ``norm``, ``mel_scale``, ``window_fn``.
    """

    def __init__(self, sr, n_fft, hop_length, f_min: float = 0.0, f_max=None, n_mels: int = 128, win_length=None,
                 power: float = 2.0, normalized: bool = False, norm=None, mel_scale: str = "htk",
                 window_fn=torch.hann_window, wkwargs=None, **unsupported) -> None:
        super().__init__()
        if unsupported:
            raise TypeError(f"MelSpec: unsupported MelSpectrogram arguments {sorted(unsupported)}")
        if normalized:
            raise NotImplementedError("MelSpec: normalized=True is not built")
        if power not in (1, 1.0, 2, 2.0):
            raise NotImplementedError("MelSpec: power must be 1 or 2")
        if n_fft & (n_fft - 1) or not 128 <= n_fft <= 4096:
            raise NotImplementedError("MelSpec: n_fft must be a power of two in [128, 4096]")
        self.sample_rate, self.n_fft, self.hop_length = int(sr), int(n_fft), int(hop_length)
        self.power = float(power)
        self.n_mels = int(n_mels)
        win_length = int(win_length) if win_length is not None else self.n_fft
        window = window_fn(win_length, **(wkwargs or {})).float()
        f_max = float(f_max) if f_max is not None else float(self.sample_rate // 2)
        fb = mel_filterbank(self.n_fft // 2 + 1, float(f_min), f_max, self.n_mels, self.sample_rate, norm, mel_scale)
        self.mel = nn.Sequential(
            _ReflectionPadSpec((self.n_fft // 2 - self.hop_length // 2, self.n_fft // 2 + self.hop_length // 2)),
            _MelSpectrogramState(window, fb))
        self._cache = None

    def _device_tables(self, device):
        """(full-length window, fb transposed to (n_mels, n_freqs), lo, hi) on `device`; rebuilt when the buffers change identity or version."""
        st = self.mel[1]
        window, fb = st.spectrogram.window, st.mel_scale.fb
        key = (window.data_ptr(), window._version, fb.data_ptr(), fb._version, str(device))
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        w = window.detach().to(device=device, dtype=torch.float32)
        if w.numel() < self.n_fft:                              # torch.stft centres a short window in the frame
            left = (self.n_fft - w.numel()) // 2
            w = torch.nn.functional.pad(w, (left, self.n_fft - w.numel() - left))
        f = fb.detach().to(device=device, dtype=torch.float32).contiguous()
        nz = f != 0
        any_nz = nz.any(0)
        lo = torch.where(any_nz, nz.float().argmax(0), torch.zeros_like(any_nz, dtype=torch.long))
        hi = torch.where(any_nz, f.shape[0] - nz.flip(0).float().argmax(0), torch.zeros_like(any_nz, dtype=torch.long))
        tables = (w.contiguous(), f.t().contiguous(), lo.int().contiguous(), hi.int().contiguous())
        self._cache = (key, tables)
        return tables

    def forward(self, x: Tensor) -> Tensor:
        L.require_cuda(x, op="MelSpec")
        squeeze = x.dim() == 1
        if squeeze:
            x = x.unsqueeze(0)
        if x.dim() != 2:
            raise ValueError(f"MelSpec expects (B, T) or (T,), got {tuple(x.shape)}")
        x = x.detach()
        if x.dtype != torch.float32 or x.stride(1) != 1:
            x = x.float().contiguous()
        B, T = x.shape
        lib = L.load()
        frames = lib.cmwg_melspec_frames(T, self.n_fft, self.hop_length)
        window, fbt, lo, hi = self._device_tables(x.device)
        out = torch.empty((B, self.n_mels, frames), device=x.device, dtype=torch.float32)
        L.check(lib.cmwg_melspec_fwd(x.data_ptr(), x.stride(0) if B > 1 else T, B, T, window.data_ptr(), fbt.data_ptr(),
                                     lo.data_ptr(), hi.data_ptr(), self.n_fft, self.hop_length, self.n_mels,
                                     int(self.power == 1.0), 1e-7, 1, out.data_ptr(), L.stream_ptr(x.device)),
                "melspec_fwd")
        return out[0] if squeeze else out
