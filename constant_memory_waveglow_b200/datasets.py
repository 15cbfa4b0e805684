"""Data path in front of the flow: wave-file segments -> pinned host batches -> device.

``RandomWAVDataset(data_dir, size, segment, deterministic=True)`` keeps the constructor, the length and the
crop-index arithmetic of the reference's ``datasets/random_wav.py:16-65``: every file contributes
``max(0, frames - segment) + 1`` start positions, the positions of all files are laid end to end on [0, 1) and item
``index`` reads the segment at the position ``index / size`` (or a uniform random one).  The reference reads audio
through ``torchaudio.info`` / ``torchaudio.load``; torchaudio 2.11 has neither without the absent TorchCodec, so
the RIFF/WAVE container is parsed here (``wav_info`` / ``wav_read``: PCM 8/16/24/32 bit and IEEE float 32/64,
plain and WAVE_FORMAT_EXTENSIBLE headers), scaled like torchaudio's ``normalize=True`` (int / 2^(bits-1)).

``DevicePrefetcher`` overlaps the host->device copy of batch k+1 with the training step of batch k: batches are
staged in two pinned buffers and copied on a side stream; the consumer waits on an event, never on the host.
"""
from __future__ import annotations

import os
import random
import struct
from pathlib import Path
from typing import Iterable, Iterator, NamedTuple, Optional

import numpy as np
import torch
from torch.utils.data import Dataset

__all__ = ["RandomWAVDataset", "DevicePrefetcher", "wav_info", "wav_read", "wav_write", "WavInfo"]

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


class WavInfo(NamedTuple):
    sample_rate: int
    num_frames: int
    num_channels: int
    bits_per_sample: int
    fmt: int            # 1 = integer PCM, 3 = IEEE float
    data_offset: int    # byte offset of the first frame


def wav_info(path) -> WavInfo:
    """Header of a RIFF/WAVE file (what the reference asks ``torchaudio.info`` for, ``datasets/random_wav.py:35``)."""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = f.read(size + (size & 1))
                tag, ch, sr, _, align, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == _EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, sr, align, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, ch, sr, align, bits = fmt
                if tag not in (_PCM, _FLOAT) or bits not in (8, 16, 24, 32, 64) or ch < 1:
                    raise ValueError(f"{path}: unsupported WAVE encoding (format {tag}, {bits} bit)")
                off = f.tell()
                avail = os.fstat(f.fileno()).st_size - off
                size = min(size, avail) if size not in (0, 0xFFFFFFFF) else avail
                return WavInfo(sr, size // (ch * bits // 8), ch, bits, tag, off)
            else:
                f.seek(size + (size & 1), 1)


def wav_read(path, frame_offset: int = 0, num_frames: int = -1, info: Optional[WavInfo] = None) -> torch.Tensor:
    """(channels, frames) float32 in [-1, 1): frames ``[frame_offset, frame_offset + num_frames)`` of the file,
    clipped to its end (``torchaudio.load(f, frame_offset=..., num_frames=...)``, ``datasets/random_wav.py:59-60``)."""
    m = info or wav_info(path)
    start = min(max(frame_offset, 0), m.num_frames)
    n = m.num_frames - start if num_frames < 0 else min(num_frames, m.num_frames - start)
    bps = m.bits_per_sample // 8
    with open(path, "rb") as f:
        f.seek(m.data_offset + start * m.num_channels * bps)
        raw = f.read(n * m.num_channels * bps)
    n = len(raw) // (m.num_channels * bps)
    raw = raw[:n * m.num_channels * bps]
    if m.fmt == _FLOAT:
        x = np.frombuffer(raw, dtype="<f4" if bps == 4 else "<f8").astype(np.float32)
    elif bps == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif bps == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif bps == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    else:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.reshape(n, m.num_channels).T))


def wav_write(path, x: torch.Tensor, sample_rate: int) -> None:
    """16-bit PCM file from (channels, frames) or (frames,) float audio (``torchaudio.save``, ``inference.py:59``)."""
    x = x.detach().float().cpu()
    if x.dim() == 1:
        x = x.unsqueeze(0)
    pcm = (x.t().clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16).contiguous().numpy()
    ch, n = x.shape
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + n * ch * 2) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, _PCM, ch, sample_rate, sample_rate * ch * 2, ch * 2, 16))
        f.write(b"data" + struct.pack("<I", n * ch * 2))
        f.write(pcm.astype("<i2").tobytes())


class RandomWAVDataset(Dataset):
    """Reference ``datasets/random_wav.py:11-65``: fixed-length mono segments drawn from a directory of wave files."""

    def __init__(self, data_dir: str, size: int, segment: int, deterministic: bool = True):
        self.segment = segment
        self.data_path = os.path.expanduser(data_dir)
        self.size = size
        self.deterministic = deterministic
        self.sr = None
        self.files = []
        self._infos = []
        starts = []
        for filename in sorted(Path(self.data_path).glob("**/*.wav")):
            meta = wav_info(filename)
            self.files.append(filename)
            self._infos.append(meta)
            starts.append(max(0, meta.num_frames - segment) + 1)
            if not self.sr:
                self.sr = meta.sample_rate
            else:
                assert meta.sample_rate == self.sr
        if not self.files:
            raise FileNotFoundError(f"RandomWAVDataset: no .wav files under {self.data_path}")
        self.file_lengths = np.array(starts)
        self.boundaries = np.cumsum(np.array([0] + starts)) / self.file_lengths.sum()

    def __len__(self):
        return self.size

    def locate(self, uniform_pos: float):
        """(file index, first frame) of the segment at position `uniform_pos` in [0, 1) (``random_wav.py:52-58``)."""
        k = int(np.digitize(uniform_pos, self.boundaries[1:], right=False))
        lo, hi = self.boundaries[k], self.boundaries[k + 1]
        return k, int(self.file_lengths[k] * (uniform_pos - lo) / (hi - lo))

    def __getitem__(self, index):
        pos = index / self.size if self.deterministic else random.uniform(0, 1)
        k, offset = self.locate(pos)
        x = wav_read(self.files[k], offset, self.segment, self._infos[k]).mean(0)
        if x.numel() < self.segment:
            x = torch.cat([x, x.new_zeros(self.segment - x.numel())])
        return x


class DevicePrefetcher:
    """Iterate a loader of CPU tensors as device tensors, one batch ahead.

    Batch k+1 is copied into one of two pinned staging buffers and sent on a side stream while the consumer works on
    batch k; ``__next__`` makes the consumer's stream wait on the copy's event (no host synchronisation)."""

    def __init__(self, loader: Iterable, device: torch.device):
        self.loader = loader
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._pinned = [None, None]
        self._copied = [None, None]   # event of the last copy OUT of each pinned buffer
        self._slot = 0

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch: torch.Tensor):
        if not self.cuda:
            return batch.to(self.device), None
        slot = self._slot
        self._slot ^= 1
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()   # the buffer's previous contents have left the host
        buf = self._pinned[slot]
        if buf is None or buf.shape != batch.shape or buf.dtype != batch.dtype:
            buf = self._pinned[slot] = torch.empty(batch.shape, dtype=batch.dtype).pin_memory()
        buf.copy_(batch)
        with torch.cuda.stream(self.stream):
            dev = buf.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        return dev, ev

    def __iter__(self) -> Iterator[torch.Tensor]:
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                cur.record_stream(torch.cuda.current_stream(self.device))
            yield cur
