"""Model-level parity the reference's own tests lack: WaveGlow against the golden fixture of the unmodified reference and
against the CPU oracle, per precision mode.  The block-level properties (input storage freed and restored, exact
invertibility, naive vs constant-memory gradients) are checked by the reference's OWN tests/test_fwd_bwd.py, run unmodified
against these modules by tests/test_reference_suite.py -- no transcription of it lives here."""
import pytest
import torch

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from oracle import flow_oracle as O
from tests._util import TOL, load_golden, rel_l2, to_double

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_mode():
    # tests/test_fwd_bwd.py:10-11 of the reference
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, precision.get_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    precision.set_precision("auto")
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old[0], old[1]
    precision.set_precision(old[2])


# ------------------------------------------------------------------------------------------------
# model level
# ------------------------------------------------------------------------------------------------
def test_waveglow_against_reference_golden():
    fx = load_golden("waveglow_tiny.pt")
    m = cm.WaveGlow(memory_efficient=True, **fx["arch"], **fx["wn_kwargs"])
    m.load_state_dict(fx["state"])
    m = m.cuda().train()
    x, h = fx["x"].cuda(), fx["h"].cuda()
    z, logdet = m(x.clone(), h)
    assert rel_l2(z, fx["z"]) < 5e-6, rel_l2(z, fx["z"])
    assert rel_l2(logdet, fx["logdet"]) < 5e-5
    loss = cm.WaveGlowLoss(fx["sigma"])(z, logdet)
    assert abs(loss.item() - fx["loss"].item()) < 1e-5 * abs(fx["loss"].item())
    loss.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert rel_l2(p.grad, fx["grads"][n]) < 2e-4, (n, rel_l2(p.grad, fx["grads"][n]))
    with torch.no_grad():
        xr, ldr = m.reverse(z.detach().clone(), h)
        assert torch.allclose(xr.cpu(), fx["x"], atol=2e-5)
        assert rel_l2(ldr, fx["logdet_reverse"]) < 5e-5
        audio = m.infer(h, 0.6, z=fx["infer_z"].cuda())
        assert torch.allclose(audio.cpu(), fx["infer_audio"], atol=2e-5)
        m.apply(cm.remove_weight_norms)          # inference.py:17
        audio2 = m.infer(h, 0.6, z=fx["infer_z"].cuda())
        assert torch.allclose(audio2.cpu(), fx["infer_audio"], atol=2e-5)


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
def test_waveglow_128ch_against_oracle(prec):
    """Config 1b of SURVEY 8(d): 12 flows, 128 channels, depth 4, B=2, T=16000, sigma 0.7."""
    precision.set_precision(prec)
    spec = O.WaveGlowSpec(12, 8, 4, 2, 256, 80)
    sd = O.random_state(spec, 128, 4, seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 16000, generator=g) * 2 - 1
    h = torch.randn(2, 80, 63, generator=g)
    z_ref, ld_ref, loss_ref, grads_ref = O.waveglow_train_step(sd, spec, x, h, 0.7)
    m = cm.WaveGlow(12, 8, 4, 2, 256, 80, True, dilation_channels=128, residual_channels=128, skip_channels=128,
                    depth=4, zero_init=False)
    m.load_state_dict(sd)
    m = m.cuda().train()
    z, logdet = m(x.cuda(), h.cuda())
    loss = cm.WaveGlowLoss(0.7)(z, logdet)
    loss.backward()
    tol = TOL[prec]
    # the oracle here is fp32 (not fp64): allow its own round-off on top of ours
    assert rel_l2(z, z_ref) < max(tol["out"], 2e-5), rel_l2(z, z_ref)
    assert rel_l2(logdet, ld_ref) < max(tol["logdet"], 1e-4), rel_l2(logdet, ld_ref)
    num = den = 0.0
    for n, p in m.named_parameters():
        gr = grads_ref[n].double()
        num += (p.grad.double().cpu() - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    agg = (num / den) ** 0.5
    assert agg < max(tol["grad"], 2e-4), agg
    with torch.no_grad():
        xr, _ = m.reverse(z.detach().clone(), h.cuda())
    assert rel_l2(xr, x) < max(tol["roundtrip"], 2e-5), rel_l2(xr, x)
