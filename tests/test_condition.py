"""Mel-spectrogram conditioner (reference model/condition.py:7-19).

CPU: the numpy oracle and the host-side filterbank against fixtures produced by the unmodified reference.
GPU: the fused CUDA kernel (through the C ABI) against the fixtures and the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import condition_oracle as CO
from tests._util import load_golden

CASES = ["melspec_lj.pt", "melspec_melglow.pt", "melspec_small.pt"]
# log-mel of fp32 pipelines: the reference itself (torchaudio, fp32 FFT) sits ~1e-6 from the fp64 oracle
TOL_ABS = 2e-5


def _oracle_args(a):
    return dict(sr=a["sr"], n_fft=a["n_fft"], hop=a["hop_length"], f_min=a.get("f_min", 0.0), f_max=a.get("f_max"),
                n_mels=a["n_mels"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    fx = load_golden(name)
    out = CO.melspec(fx["x"].numpy(), **_oracle_args(fx["args"]))
    assert out.shape == tuple(fx["out"].shape)
    # the oracle's own fp64 filterbank differs from torchaudio's fp32 one by ~1e-5 relative
    assert np.abs(out - fx["out"].double().numpy()).max() < 5e-5
    out_fb = CO.melspec(fx["x"].numpy(), **_oracle_args(fx["args"]), fb=fx["fb"].numpy(), window=fx["window"].numpy())
    assert np.abs(out_fb - fx["out"].double().numpy()).max() < TOL_ABS
    fb = CO.mel_filterbank(fx["args"]["n_fft"] // 2 + 1, fx["args"].get("f_min", 0.0),
                           fx["args"].get("f_max") or float(fx["args"]["sr"] // 2), fx["args"]["n_mels"],
                           fx["args"]["sr"])
    assert np.abs(fb - fx["fb"].double().numpy()).max() < 1e-4
    assert np.abs(CO.hann_periodic(fx["args"]["n_fft"]) - fx["window"].double().numpy()).max() < 1e-6


@pytest.mark.parametrize("name", CASES)
def test_host_tables_match_reference(name):
    """Host logic (no GPU): filterbank, window, module layout and state-dict keys of the reference's MelSpec."""
    from constant_memory_waveglow_b200.condition import MelSpec
    fx = load_golden(name)
    m = MelSpec(**fx["args"])
    st = m.mel[1]
    assert torch.equal(st.mel_scale.fb, fx["fb"])          # same fp32 op sequence as torchaudio -> bit-exact
    assert torch.equal(st.spectrogram.window, fx["window"])
    assert sorted(m.state_dict().keys()) == fx["state_keys"] == ["mel.1.mel_scale.fb", "mel.1.spectrogram.window"]
    m.load_state_dict({"mel.1.mel_scale.fb": fx["fb"], "mel.1.spectrogram.window": fx["window"]})
    assert sorted(n for n, _ in m.named_buffers()) == fx["buffer_names"]
    a = fx["args"]
    assert m.mel[0].padding == (a["n_fft"] // 2 - a["hop_length"] // 2, a["n_fft"] // 2 + a["hop_length"] // 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(fx["x"])                                          # no CPU path


def test_model_package_exports_condition():
    import model
    from model import condition
    assert hasattr(condition, "MelSpec")
    assert model.condition is condition


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_fixture_and_oracle(name):
    from constant_memory_waveglow_b200.condition import MelSpec
    fx = load_golden(name)
    m = MelSpec(**fx["args"]).cuda()
    out = m(fx["x"].cuda())
    assert out.shape == fx["out"].shape and out.dtype == torch.float32
    ref64 = CO.melspec(fx["x"].numpy(), **_oracle_args(fx["args"]), fb=fx["fb"].numpy(), window=fx["window"].numpy())
    assert np.abs(out.double().cpu().numpy() - ref64).max() < TOL_ABS
    assert (out.cpu() - fx["out"]).abs().max().item() < TOL_ABS


@pytest.mark.gpu
def test_kernel_layouts_and_sizes():
    """Strided batch rows, 1-D input, every supported transform size, power 1, short-input error."""
    from constant_memory_waveglow_b200.condition import MelSpec
    g = torch.Generator().manual_seed(5)
    for n_fft, hop, T in ((128, 32, 700), (512, 128, 4096), (2048, 512, 5000), (4096, 1024, 9001)):
        m = MelSpec(16000, n_fft, hop, n_mels=40).cuda()
        wide = torch.rand(3, T + 37, generator=g) * 2 - 1
        x = wide[:, :T]                                       # batch stride T + 37
        out = m(x.cuda()[:, :T])
        ref = CO.melspec(x.numpy(), 16000, n_fft, hop, n_mels=40)
        assert out.shape == (3, 40, T // hop + 1)
        assert np.abs(out.double().cpu().numpy() - ref).max() < 5e-5, (n_fft, hop)
        one = m(x[1].cuda())
        assert torch.equal(one, out[1])
    m1 = MelSpec(22050, 1024, 256, n_mels=80, f_max=8000, power=1.0).cuda()
    x = torch.rand(2, 2000, generator=g) * 2 - 1
    ref = CO.melspec(x.numpy(), 22050, 1024, 256, f_max=8000.0, n_mels=80, power=1.0)
    assert np.abs(m1(x.cuda()).double().cpu().numpy() - ref).max() < 2e-5
    with pytest.raises(RuntimeError, match="reflection padding"):
        m1(torch.zeros(1, 600).cuda())


@pytest.mark.gpu
def test_kernel_full_size_properties():
    """BASELINE sizes: training batch (24 x 16000) and one 10 s utterance; frame-locality and determinism."""
    from constant_memory_waveglow_b200.condition import MelSpec
    m = MelSpec(22050, 1024, 256, f_max=8000, n_mels=80).cuda()
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(24, 16000, generator=g) * 2 - 1).cuda()
    out = m(x)
    assert out.shape == (24, 80, 63) and torch.isfinite(out).all()
    assert torch.equal(out, m(x))                              # deterministic
    # a frame depends only on its own n_fft samples: changing the tail leaves early frames bit-identical
    x2 = x.clone()
    x2[:, 8000:] = 0
    out2 = m(x2)
    last_clean = (8000 - 640) // 256 - 1
    assert torch.equal(out2[..., :last_clean], out[..., :last_clean])
    # silence -> log(1e-7)
    z = m(torch.zeros(1, 220672, device="cuda"))
    assert z.shape == (1, 80, 863)
    assert torch.allclose(z, torch.full_like(z, float(np.log(np.float32(1e-7)))))
    # spot check of the big batch against the oracle
    ref = CO.melspec(x[5:6].cpu().numpy(), 22050, 1024, 256, n_mels=80, fb=m.mel[1].mel_scale.fb.cpu().numpy())
    assert np.abs(out[5:6].double().cpu().numpy() - ref).max() < TOL_ABS
