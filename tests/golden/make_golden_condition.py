"""Golden fixtures of the mel-spectrogram conditioner, produced by the UNMODIFIED reference
``model/condition.py::MelSpec`` (which calls torchaudio) on CPU.  Authoring container only:

    python tests/golden/make_golden_condition.py
"""
import importlib.util
import os
import warnings

import torch

warnings.filterwarnings("ignore")
REF = os.environ.get("CMWG_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("_ref_condition", os.path.join(REF, "model", "condition.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = {
        # configs/waveglow_LJ_speech.json:51-60 (also waveflow / mr_waveglow)
        "melspec_lj": dict(args=dict(sr=22050, n_fft=1024, hop_length=256, f_max=8000, n_mels=80), B=2, T=3000, seed=0),
        # configs/melglow_LJ_speech.json conditioner
        "melspec_melglow": dict(args=dict(sr=22050, n_fft=1024, hop_length=256, f_min=60, f_max=7600, n_mels=80),
                                B=1, T=2048, seed=1),
        # a different transform size, odd length
        "melspec_small": dict(args=dict(sr=16000, n_fft=256, hop_length=64, n_mels=20), B=3, T=1001, seed=2),
    }
    for name, c in cases.items():
        g = torch.Generator().manual_seed(c["seed"])
        x = torch.rand(c["B"], c["T"], generator=g) * 2 - 1
        x[0, : c["T"] // 4] *= 1e-3                      # a quiet stretch: small mel powers next to the 1e-7 floor
        m = ref.MelSpec(**c["args"])
        with torch.no_grad():
            out = m(x.clone())
        mel = m.mel[1]
        torch.save({"args": c["args"], "x": x, "out": out.clone(), "fb": mel.mel_scale.fb.clone(),
                    "window": mel.spectrogram.window.clone(), "state_keys": sorted(m.state_dict().keys()),
                    "buffer_names": sorted(n for n, _ in m.named_buffers())},
                   os.path.join(OUT, name + ".pt"))
        print(name, tuple(out.shape), float(out.mean()))


if __name__ == "__main__":
    main()
