"""MRWaveGlow (reference model/mr_waveglow.py): the oracle restatement against fixtures generated from
the unmodified reference (CPU), and the CUDA modules against both (GPU)."""
import pytest
import torch

from oracle import flow_oracle as O
from tests._util import TOL, load_golden, max_abs, rel_l2, to_double

FIXTURES = ["mrwaveglow_tiny.pt", "mrwaveglow_tiny_sr.pt"]


def _spec(fx):
    return O.MRSpec(**fx["arch"])


def _close(a, b, rtol):
    """rel-L2 within rtol, or both numerically zero (a weight_g gradient that vanishes analytically is 1e-12 noise)."""
    return rel_l2(a, b) < rtol or max_abs(a, b) < 1e-8


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_fixture(name):
    fx = load_golden(name)
    spec = _spec(fx)
    z, logdet, loss, grads = O.mrwaveglow_train_step(fx["state"], spec, fx["x"], fx["h"], fx["sigma"])
    assert rel_l2(z, fx["z"]) < 1e-5 and rel_l2(logdet, fx["logdet"]) < 1e-5 and abs(float(loss - fx["loss"])) < 1e-6
    assert set(grads) == set(fx["grads"])
    for k, g in fx["grads"].items():
        assert _close(grads[k], g, 2e-4), k
    with torch.no_grad():
        xr, ldr = O.mrwaveglow_reverse(fx["state"], spec, fx["z"], fx["h"])
        audio, _ = O.mrwaveglow_reverse(fx["state"], spec, fx["infer_z"], fx["h"])
    assert max_abs(xr, fx["x"]) < 5e-6 and rel_l2(ldr, fx["logdet_reverse"]) < 1e-5
    assert rel_l2(audio, fx["infer_audio"]) < 1e-5


def test_module_surface_and_state_dict_keys():
    import constant_memory_waveglow_b200 as cm
    fx = load_golden(FIXTURES[0])
    m = cm.MRWaveGlow(memory_efficient=True, **fx["arch"], **fx["wn_kwargs"])
    assert list(m.state_dict().keys()) == list(fx["state"].keys())
    assert all(m.state_dict()[k].shape == v.shape for k, v in fx["state"].items())
    m.load_state_dict(fx["state"])
    # the reference's quirk: level 1x1 convs are memory-efficient whatever the flag says (mr_waveglow.py:46)
    m2 = cm.MRWaveGlow(memory_efficient=False, **fx["arch"], **fx["wn_kwargs"])
    assert hasattr(m2.invconv1x1_list[0][0], "_efficient_forward") and not hasattr(m2.prior_invconv1x1[0], "_efficient_forward")
    from model import MRWaveGlow  # the drop-in import name (reference model/__init__.py)
    assert MRWaveGlow is cm.MRWaveGlow
    with pytest.raises(RuntimeError):
        m(fx["x"].clone(), fx["h"])          # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("efficient", [True, False])
def test_gpu_forward_backward_reverse_against_fixture(name, efficient):
    """fp32 engine (16-channel WN): z, logdet, loss, every gradient, the round trip and synthesis from fixed noise."""
    import constant_memory_waveglow_b200 as cm
    fx = load_golden(name)
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = cm.MRWaveGlow(memory_efficient=efficient, **fx["arch"], **fx["wn_kwargs"]).cuda()
        m.load_state_dict(fx["state"])
        loss_fn = cm.WaveGlowLoss(fx["sigma"])
        z, logdet = m(fx["x"].cuda().clone(), fx["h"].cuda())
        loss = loss_fn(z, logdet)
        loss.backward()
        tol = TOL["fp32"]
        assert rel_l2(z, fx["z"]) < 10 * tol["out"] and rel_l2(logdet, fx["logdet"]) < tol["logdet"]
        assert abs(float(loss) - float(fx["loss"])) < 1e-5
        for n, p in m.named_parameters():
            assert p.grad is not None and _close(p.grad, fx["grads"][n], 2e-4), n
        with torch.no_grad():
            xr, ldr = m.reverse(z.detach().clone(), fx["h"].cuda())
            audio = m.infer(fx["h"].cuda(), 0.6, z=fx["infer_z"].cuda())
        assert max_abs(xr, fx["x"]) < 1e-5 and rel_l2(ldr, fx["logdet_reverse"]) < tol["logdet"]
        assert rel_l2(audio, fx["infer_audio"].squeeze()) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = True


@pytest.mark.gpu
def test_gpu_lj_width_bf16_against_oracle():
    """The shipped width (256 channels: tcgen05 engine, task kernels) at a small size against the fp64 oracle,
    including the gradient through the signal-dependent conditioning of the level flows."""
    import constant_memory_waveglow_b200 as cm
    from constant_memory_waveglow_b200 import precision
    torch.manual_seed(4)
    arch = dict(prior_flows=2, n_group=8, hop_size=256, n_mels=80, levels=3, flows=1)
    wkw = dict(dilation_channels=256, residual_channels=256, skip_channels=256, depth=3, radix=3, bias=False, zero_init=False)
    m = cm.MRWaveGlow(memory_efficient=True, **arch, **wkw).cuda()
    x = torch.rand(2, 4096) * 2 - 1
    h = torch.randn(2, 80, 16)
    sd = to_double({k: v.cpu() for k, v in m.state_dict().items()})
    zo, ldo, losso, go = O.mrwaveglow_train_step(sd, O.MRSpec(**arch), x.double(), h.double(), 0.7)
    old = precision.get_precision()
    precision.set_precision("fp16")
    try:
        xin = x.cuda().requires_grad_(True)
        z, logdet = m(xin * 1.0, h.cuda())
        loss = cm.WaveGlowLoss(0.7)(z, logdet)
        loss.backward()
    finally:
        precision.set_precision(old)
    tol = TOL["fp16"]
    assert rel_l2(z, zo) < tol["out"] and rel_l2(logdet, ldo) < tol["logdet"]
    for n, p in m.named_parameters():
        assert _close(p.grad, go[n], tol["grad_worst"]), n
    assert torch.isfinite(xin.grad).all()
