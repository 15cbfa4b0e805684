"""Module-level parity: this file reads like the reference's tests/test_fwd_bwd.py (same grids, same
five properties per test, TF32 disabled -> exact engine), plus the model-level checks the reference
lacks: WaveGlow against the golden fixture of the unmodified reference and against the fp64 oracle."""
import numpy as np
import pytest
import torch
from torch import nn

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from model.efficient_modules import AffineCouplingBlock, InvertibleConv1x1
from model.loss import WaveGlowLoss
from model.waveglow import WN
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


def set_seed(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)


def storage_freed(t):
    return t.untyped_storage().size() == 0


@pytest.mark.parametrize('batch', [1, 4, 32])
@pytest.mark.parametrize('channels', [2, 4, 8])
@pytest.mark.parametrize('length', [2000])
def test_conv1x1_fwd_bwd(batch, channels, length):
    weights = InvertibleConv1x1(channels).state_dict()
    loss_func = WaveGlowLoss().cuda()
    for seed in range(3):
        set_seed(seed)
        data = torch.rand(batch, channels, length) * 2 - 1
        for bwd in [False, True]:
            impl_out, impl_grad = [], []
            for keep_input in [True, False]:
                model = InvertibleConv1x1(channels, not keep_input)
                model.load_state_dict(weights)
                model = model.cuda()
                model.train()
                model.zero_grad()
                x = data.cuda()
                xin = x.clone()
                if bwd:
                    y, log1 = model.reverse(xin)
                    yrev = y.clone()
                    xinv, log2 = model(yrev)
                else:
                    y, log1 = model(xin)
                    yrev = y.clone()
                    xinv, log2 = model.reverse(yrev)
                assert torch.equal(log1, log2.neg())
                assert log1.dim() == 0
                loss = loss_func(y.view(batch, -1), log1)
                if keep_input:
                    assert xin.shape == x.shape and not storage_freed(xin)
                else:
                    assert storage_freed(xin) and storage_freed(yrev)
                loss.backward()
                assert y.shape == x.shape
                assert torch.allclose(x.cpu(), data)
                assert torch.allclose(x, xinv, atol=1e-6, rtol=0)
                if not keep_input:
                    assert not storage_freed(xin) and torch.allclose(xin, x, atol=1e-6, rtol=0)
                impl_out.append(y.detach().cpu())
                impl_grad.append([p.grad.cpu() for p in model.parameters()])
            for g1, g2 in zip(impl_grad[0], impl_grad[1]):
                assert torch.allclose(g1, g2, atol=5e-7, rtol=0)
            assert torch.allclose(impl_out[0], impl_out[1])


@pytest.mark.parametrize('batch', [2])
@pytest.mark.parametrize('channels', [16, 32])
@pytest.mark.parametrize('WN_channels', [128])
@pytest.mark.parametrize('depth', [1, 4])
@pytest.mark.parametrize('aux_channels', [20, 40])
@pytest.mark.parametrize('length', [4000])
def test_affine_fwd_bwd(batch, channels, WN_channels, depth, aux_channels, length):
    kw = dict(in_channels=channels // 2, aux_channels=aux_channels, zero_init=False, dilation_channels=WN_channels,
              residual_channels=WN_channels, skip_channels=WN_channels, depth=depth)
    weights = AffineCouplingBlock(WN, False, **kw).state_dict()
    loss_func = WaveGlowLoss().cuda()
    for seed in range(2):
        set_seed(seed)
        data = torch.rand(batch, channels, length) * 2 - 1
        condition = torch.randn(batch, aux_channels, length)
        for bwd in [False, True]:
            impl_out, impl_grad = [], []
            for keep_input in [True, False]:
                model = AffineCouplingBlock(WN, not keep_input, **kw)
                model.load_state_dict(weights)
                model = model.cuda()
                model.train()
                model.zero_grad()
                x = data.cuda()
                h = condition.cuda()
                xin = x.clone()
                if bwd:
                    y, log1 = model.reverse(xin, h)
                    yrev = y.clone()
                    xinv, log2 = model(yrev, h)
                else:
                    y, log1 = model(xin, h)
                    yrev = y.clone()
                    xinv, log2 = model.reverse(yrev, h)
                assert torch.equal(log1, log2.neg())
                loss = loss_func(y.view(2, -1), log1.sum((1, 2)))
                if keep_input:
                    assert not storage_freed(xin)
                else:
                    assert storage_freed(xin) and storage_freed(yrev)
                    assert torch.allclose(h.cpu(), condition)
                loss.backward()
                assert torch.allclose(x.cpu(), data)
                assert torch.allclose(x, xinv, atol=1e-6)
                impl_out.append(y.cpu().detach())
                impl_grad.append([p.grad.cpu() for p in model.parameters()])
            for g1, g2 in zip(impl_grad[0], impl_grad[1]):
                # same kernels on bit-identical recomputed activations; only the restored xb carries
                # fp32 round-off into the log_s cotangent
                assert torch.allclose(g1, g2, rtol=1e-4, atol=1e-7)
            assert torch.allclose(impl_out[0], impl_out[1])


@pytest.mark.parametrize('batch', [2, 16])
@pytest.mark.parametrize('channels', [2, 8])
@pytest.mark.parametrize('length', [2000])
def test_complx_chained(batch, channels, length):
    model1 = nn.ModuleList([InvertibleConv1x1(channels, True), InvertibleConv1x1(channels, False),
                            InvertibleConv1x1(channels, True)])
    model2 = nn.ModuleList([InvertibleConv1x1(channels, False), InvertibleConv1x1(channels, True),
                            InvertibleConv1x1(channels, False)])
    model2.load_state_dict(model1.state_dict())
    loss_func = WaveGlowLoss().cuda()
    for seed in range(3):
        set_seed(seed)
        data = torch.rand(batch, channels, length) * 2 - 1
        impl_grad = []
        for model in [model1, model2]:
            model = model.cuda()
            model.train()
            model.zero_grad()
            xin = data.cuda().clone()
            logdet = 0
            for layer in model:
                xin, ld = layer.reverse(xin)
                logdet = logdet + ld
            loss = loss_func(xin.view(batch, -1), logdet)
            loss.backward()
            impl_grad.append([p.grad.cpu() for p in model.parameters()])
        for g1, g2 in zip(impl_grad[0], impl_grad[1]):
            assert torch.allclose(g1, g2, atol=5e-7, rtol=0)


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
