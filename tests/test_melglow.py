"""MelGlow (reference model/melglow.py): oracle restatement and the WN_LVC transform against a fixture of the
unmodified reference (CPU); the flow on the device against the fixture (GPU)."""
import pytest
import torch

from oracle import flow_oracle as O
from tests._util import load_golden, max_abs, rel_l2


def _close(a, b, rtol):
    return rel_l2(a, b) < rtol or max_abs(a, b) < 1e-8


def _spec(fx):
    kw = {k: v for k, v in fx["wn_kwargs"].items() if k not in ("bias", "zero_init")}
    return O.MelGlowSpec(**fx["arch"], **kw)


def test_oracle_matches_reference_fixture():
    fx = load_golden("melglow_tiny.pt")
    spec = _spec(fx)
    with torch.no_grad():
        log_s, t = O.wn_lvc_forward(fx["state"], "WNs.0.F.", spec, fx["wn_x"], fx["h"][..., :16])
    assert rel_l2(log_s, fx["wn_log_s"]) < 1e-5 and rel_l2(t, fx["wn_t"]) < 1e-5
    z, logdet, loss, grads = O.melglow_train_step(fx["state"], spec, fx["x"], fx["h"], fx["sigma"])
    assert rel_l2(z, fx["z"]) < 1e-5 and rel_l2(logdet, fx["logdet"]) < 1e-5 and abs(float(loss - fx["loss"])) < 1e-6
    assert set(grads) == set(fx["grads"])
    for k, g in fx["grads"].items():
        assert _close(grads[k], g, 5e-4), k
    with torch.no_grad():
        xr, ldr = O.melglow_reverse(fx["state"], spec, fx["z"], fx["h"])
    assert max_abs(xr, fx["x"]) < 5e-6 and rel_l2(ldr, fx["logdet_reverse"]) < 1e-5


def test_wn_lvc_module_layout_matches_reference_fixture():
    """Same state-dict keys and parameter order as the reference; no CPU path (the location-variable convolutions run in
    csrc/lvc.cu like the flow primitives)."""
    import constant_memory_waveglow_b200 as cm
    fx = load_golden("melglow_tiny.pt")
    m = cm.MelGlow(memory_efficient=True, **fx["arch"], **fx["wn_kwargs"]).train()
    assert list(m.state_dict().keys()) == list(fx["state"].keys())
    m.load_state_dict(fx["state"])
    wn = m.WNs[0].F
    assert [n for n, _ in wn.named_parameters()][0].startswith("start.")
    from model import MelGlow
    assert MelGlow is cm.MelGlow
    with pytest.raises(RuntimeError):
        wn(fx["wn_x"], fx["h"][..., :16])
    with pytest.raises(RuntimeError):
        m(fx["x"].clone(), fx["h"])          # the flow primitives have no CPU path


@pytest.mark.gpu
def test_wn_lvc_kernels_against_fixture_and_torch_ops():
    """cmwg_lvc_gate_forward / _backward: the WN_LVC transform against the reference fixture, and one layer's forward,
    input gradient and kernel gradient against the same computation written as torch ops -- dilations 1 .. 64 put the
    taps' source samples in the same, the neighbouring and the second-next frame."""
    import constant_memory_waveglow_b200 as cm
    from constant_memory_waveglow_b200.melglow import NonCausalLayerLVC, _LVCGate
    fx = load_golden("melglow_tiny.pt")
    m = cm.MelGlow(memory_efficient=True, **fx["arch"], **fx["wn_kwargs"]).train()
    m.load_state_dict(fx["state"])
    wn = m.WNs[0].F.cuda()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the predictor / W_o are torch convs
    try:
        with torch.no_grad():
            log_s, t = wn(fx["wn_x"].cuda(), fx["h"][..., :16].cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert rel_l2(log_s, fx["wn_log_s"]) < 1e-5 and rel_l2(t, fx["wn_t"]) < 1e-5
    g = torch.Generator().manual_seed(0)
    for (B, frames, span, Cd, Cr, radix, dil) in ((2, 5, 32, 48, 48, 3, 1), (3, 7, 32, 48, 48, 3, 32), (2, 6, 32, 48, 48, 3, 64),
                                                   (2, 4, 16, 8, 12, 5, 3), (1, 3, 64, 16, 8, 3, 16)):
        layer = NonCausalLayerLVC(dil, Cd, Cr, Cd, radix, False)
        x = torch.randn(B, Cr, frames * span, generator=g).cuda().requires_grad_(True)
        w = (torch.randn(B, frames, 2 * Cd, Cr, radix, generator=g) * 0.2).cuda().requires_grad_(True)
        dg = torch.randn(B, Cd, frames * span, generator=g).cuda()
        want = layer._lvc_gate_torch(x, w)
        gx, gw = torch.autograd.grad(want, [x, w], dg)
        got = _LVCGate.apply(x, w, dil)
        hx, hw = torch.autograd.grad(got, [x, w], dg)
        assert rel_l2(got, want) < 2e-6, (dil, rel_l2(got, want))
        assert rel_l2(hx, gx) < 5e-6 and rel_l2(hw, gw) < 5e-6, (dil, rel_l2(hx, gx), rel_l2(hw, gw))
        assert torch.equal(hw, torch.autograd.grad(_LVCGate.apply(x, w, dil), [w], dg)[0])     # deterministic


@pytest.mark.gpu
@pytest.mark.parametrize("efficient", [True, False])
def test_gpu_flow_against_fixture(efficient):
    import constant_memory_waveglow_b200 as cm
    fx = load_golden("melglow_tiny.pt")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = cm.MelGlow(memory_efficient=efficient, **fx["arch"], **fx["wn_kwargs"]).cuda().train()
        m.load_state_dict(fx["state"])
        z, logdet = m(fx["x"].cuda().clone(), fx["h"].cuda())
        loss = cm.WaveGlowLoss(fx["sigma"])(z, logdet)
        loss.backward()
        assert rel_l2(z, fx["z"]) < 2e-5 and rel_l2(logdet, fx["logdet"]) < 2e-5
        assert abs(float(loss) - float(fx["loss"])) < 1e-5
        for n, p in m.named_parameters():
            assert p.grad is not None and _close(p.grad, fx["grads"][n], 1e-3), n
        if efficient:   # forward + recompute moved the BatchNorm statistics exactly as in the reference
            sd = m.state_dict()
            for k, v in fx["bn_after"].items():
                assert torch.allclose(sd[k].cpu().float(), v.float(), rtol=1e-4, atol=1e-6), k
        m.load_state_dict(fx["state"])
        with torch.no_grad():
            xr, ldr = m.reverse(fx["z"].cuda().clone(), fx["h"].cuda())
        assert max_abs(xr, fx["x"]) < 2e-5 and rel_l2(ldr, fx["logdet_reverse"]) < 2e-5
    finally:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
