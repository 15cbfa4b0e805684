"""Parity at the configuration and precision bench.py times: BASELINE.json config 2 (WaveGlow LJ: 12 flows, 256 channels,
8 WN layers, 80 mel, 16000-sample segments) through the task kernels (wn_fwd_mega_kernel / wn_bwd_mega_kernel /
tc_wgrad_kernel), against the CPU oracle on the same seeded inputs and weights.

Tolerances are north_star's: rel-L2 <= 1e-3 for the default tensor-core mode (fp16 operands, fp32 accumulate) on z, log-det,
synthesis audio, the round trip and the parameter gradients as one vector (worst single tensor: 3e-3), and <= 1e-5 for the
exact fp32 engine."""
import ctypes as C

import pytest
import torch

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import _lib, precision
from oracle import flow_oracle as O
from tests._util import TOL, grad_errors, rel_l2

pytestmark = pytest.mark.gpu

B, T, FRAMES, SIGMA = 2, 16000, 63, 0.7


@pytest.fixture(autouse=True)
def _restore_precision():
    old = precision.get_precision()
    yield
    precision.set_precision(old)


@pytest.fixture(scope="module")
def lj():
    """Oracle results of one training step + one synthesis call at the LJ config (a few seconds of CPU work, once)."""
    spec = O.WaveGlowSpec(12, 8, 4, 2, 256, 80)
    sd = O.random_state(spec, 256, 8, seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, T, generator=g) * 2 - 1
    h = torch.randn(B, 80, FRAMES, generator=g)
    zs = torch.randn(B, FRAMES * 256, generator=g) * 0.6
    z, ld, loss, grads = O.waveglow_train_step(sd, spec, x, h, SIGMA)
    audio = O.waveglow_infer(sd, spec, h, zs)
    return dict(sd=sd, x=x, h=h, zs=zs, z=z, ld=ld, loss=loss, grads=grads, audio=audio)


def _kernel_classes(fn):
    lib = _lib.load()
    lib.cmwg_profile_enable(1)
    out = fn()
    torch.cuda.synchronize()
    ms, n = (C.c_double * 8)(), (C.c_longlong * 8)()
    _lib.check(lib.cmwg_profile_collect(ms, n), "profile_collect")
    lib.cmwg_profile_enable(0)
    names = ["gate", "resskip", "dgate", "dx", "dcond", "wgrad", "fwdfused", "bwdfused"]
    return out, {k: int(n[i]) for i, k in enumerate(names)}


@pytest.mark.parametrize("prec", ["fp16", "fp32"])
def test_waveglow_lj_against_oracle(lj, prec):
    precision.set_precision(prec)
    tol = TOL[prec]
    m = cm.WaveGlow(12, 8, 4, 2, 256, 80, True, zero_init=False)
    m.load_state_dict(lj["sd"])
    m = m.cuda().train()

    def train_step():
        z, logdet = m(lj["x"].cuda(), lj["h"].cuda())
        loss = cm.WaveGlowLoss(SIGMA)(z, logdet)
        loss.backward()
        return z, logdet, loss

    (z, logdet, loss), launched = _kernel_classes(train_step)
    if prec == "fp16":
        # the benchmarked path: one forward task kernel per flow (+ one per recompute), one backward chain per flow
        assert launched["fwdfused"] == 24 and launched["bwdfused"] == 12 and launched["wgrad"] >= 12, launched
        assert launched["gate"] == 0 and launched["dgate"] == 0, launched
    # the oracle here is fp32: its own round-off (~5e-7) is allowed on top of the fp32 engine's
    assert rel_l2(z, lj["z"]) < max(tol["out"], 5e-6), rel_l2(z, lj["z"])
    assert rel_l2(logdet, lj["ld"]) < max(tol["logdet"], 2e-5), rel_l2(logdet, lj["ld"])
    assert abs(loss.item() - lj["loss"].item()) < max(tol["out"], 1e-5) * abs(lj["loss"].item())
    agg, worst, name = grad_errors([(n, p.grad) for n, p in m.named_parameters()], lj["grads"])
    assert agg < max(tol["grad"], 2e-5), ("gradient aggregate", agg)
    assert worst < max(tol["grad_worst"], 1e-4), (name, worst)
    with torch.no_grad():
        m.eval()
        xr, ldr = m.reverse(z.detach().clone(), lj["h"].cuda())
        assert rel_l2(xr, lj["x"]) < max(tol["roundtrip"], 5e-6), ("round trip", rel_l2(xr, lj["x"]))
        assert rel_l2(ldr, -lj["ld"]) < max(tol["logdet"], 2e-5)
        audio = m.infer(lj["h"].cuda(), 0.6, z=lj["zs"].cuda())
        assert rel_l2(audio, lj["audio"]) < max(tol["out"], 5e-6), ("audio", rel_l2(audio, lj["audio"]))


def test_waveglow_lj_small_batch_strong_scaling_shape(lj):
    """3 segments per GPU -- the reference's split of its global batch 24 over 8 GPUs (train.py:51-53): 24 row tiles for 74
    CTA pairs, so most pairs idle or run ahead of their dependencies.  Same tolerances as the full batch."""
    precision.set_precision("fp16")
    spec = O.WaveGlowSpec(12, 8, 4, 2, 256, 80)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(3, T, generator=g) * 2 - 1
    h = torch.randn(3, 80, FRAMES, generator=g)
    z_ref, ld_ref, loss_ref, grads_ref = O.waveglow_train_step(lj["sd"], spec, x, h, SIGMA)
    m = cm.WaveGlow(12, 8, 4, 2, 256, 80, True, zero_init=False)
    m.load_state_dict(lj["sd"])
    m = m.cuda().train()
    z, logdet = m(x.cuda(), h.cuda())
    cm.WaveGlowLoss(SIGMA)(z, logdet).backward()
    tol = TOL["fp16"]
    assert rel_l2(z, z_ref) < tol["out"] and rel_l2(logdet, ld_ref) < tol["logdet"]
    agg, worst, name = grad_errors([(n, p.grad) for n, p in m.named_parameters()], grads_ref)
    assert agg < tol["grad"] and worst < tol["grad_worst"], (agg, name, worst)
