"""Fixture of the data path, generated from the UNMODIFIED reference dataset class.

Run here (needs /root/reference):  python tests/golden/make_golden_data.py
The reference's ``datasets/random_wav.py`` reads audio through ``torchaudio.info`` / ``torchaudio.load``, which
torchaudio 2.11 no longer has without TorchCodec; both are stubbed IN MEMORY with the standard library's ``wave``
module (16-bit PCM, value / 32768 like torchaudio's normalisation).  The index arithmetic under test is the reference's own.
Stored: the PCM of six small files (mono and stereo, one shorter than a segment) and, for every index of two dataset
sizes, the file the reference opened, the frame offset it asked for and the segment it returned.
"""
import importlib.util
import os
import sys
import tempfile
import types
import wave

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "random_wav.pt")
SEGMENT = 1000
FILES = [("a/one.wav", 3217, 1), ("a/two.wav", 1000, 1), ("b/three.wav", 640, 1), ("b/four.wav", 5003, 2),
         ("five.wav", 1001, 1), ("six.wav", 2500, 2)]


def write_files(root, pcm):
    for (name, frames, ch), data in zip(FILES, pcm):
        path = os.path.join(root, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with wave.open(path, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(22050)
            w.writeframes(data.astype("<i2").tobytes())


def main():
    rng = np.random.default_rng(7)
    pcm = [rng.integers(-20000, 20000, size=(frames, ch), dtype=np.int16) for _, frames, ch in FILES]
    calls = []

    ta = types.ModuleType("torchaudio")

    def info(path):
        with wave.open(str(path), "rb") as w:
            return types.SimpleNamespace(num_frames=w.getnframes(), sample_rate=w.getframerate())

    def load(path, frame_offset=0, num_frames=-1):
        with wave.open(str(path), "rb") as w:
            ch = w.getnchannels()
            w.setpos(min(frame_offset, w.getnframes()))
            raw = w.readframes(w.getnframes() if num_frames < 0 else num_frames)
            sr = w.getframerate()
        x = np.frombuffer(raw, dtype="<i2").reshape(-1, ch).T.astype(np.float32) / 32768.0
        calls.append((str(path), int(frame_offset)))
        return torch.from_numpy(np.ascontiguousarray(x)), sr

    ta.info, ta.load = info, load
    sys.modules["torchaudio"] = ta
    spec = importlib.util.spec_from_file_location("ref_random_wav", os.path.join(REF, "datasets", "random_wav.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    cases = {}
    with tempfile.TemporaryDirectory() as root:
        write_files(root, pcm)
        for size in (37, 256):
            ds = mod.RandomWAVDataset(root, size, SEGMENT)
            files, offsets, segs = [], [], []
            for i in range(size):
                calls.clear()
                x = ds[i]
                files.append(os.path.relpath(calls[0][0], root))
                offsets.append(calls[0][1])
                segs.append(x)
            # segments as exact integers: value * 65536 (the mean of two 16-bit channels is a multiple of 2^-16)
            seg = (torch.stack(segs).double() * 65536).round().int() if size == 37 else None
            cases[size] = dict(files=files, offsets=offsets, segments_x65536=seg,
                               boundaries=torch.from_numpy(ds.boundaries.copy()), sr=ds.sr)
            assert len(ds) == size
    torch.save(dict(files=FILES, pcm=[torch.from_numpy(p) for p in pcm], segment=SEGMENT, cases=cases), OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
