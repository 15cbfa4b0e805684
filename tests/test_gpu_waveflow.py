"""WaveFlow (reference model/waveflow.py) on the CUDA path against the CPU oracle and the fixtures generated from
the unmodified reference: conditioning upsampler, affine kernels, the 2-D WN (forward, backward, row-recurrent
line windows) and the model (forward + loss + backward, synthesis direction, round trip)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import _lib as L
from constant_memory_waveglow_b200 import precision
from constant_memory_waveglow_b200.waveflow import _affine, _DenseUpsampleFunction
from oracle import flow_oracle as O
from tests._util import TOL, load_golden, rel_l2, to_double

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    old = precision.get_precision()
    yield
    precision.set_precision(old)


def _wkw(ch):
    return dict(dilation_channels=ch, residual_channels=ch, skip_channels=ch, bias=False, zero_init=False)


@pytest.mark.parametrize("n_mels,n_group,frames,B", [(8, 32, 3, 2), (80, 64, 7, 3), (5, 16, 1, 1)])
def test_dense_upsampler_against_oracle(n_mels, n_group, frames, B):
    spec = O.WaveFlowSpec(1, n_group, n_mels)
    sd = O.waveflow_random_state(spec, 4, seed=n_group)
    g = torch.Generator().manual_seed(1)
    h = torch.randn(B, n_mels, frames, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("upsampler")}
    y_ref = O.waveflow_upsample_h(leaf, spec, h)
    w = torch.randn(y_ref.shape, generator=g)
    grads_ref = torch.autograd.grad((y_ref * w).sum(), [leaf["upsampler.1.weight_g"], leaf["upsampler.1.weight_v"],
                                                        leaf["upsampler.1.bias"]])
    gg = sd["upsampler.1.weight_g"].cuda().requires_grad_(True)
    vv = sd["upsampler.1.weight_v"].cuda().requires_grad_(True)
    bb = sd["upsampler.1.bias"].cuda().requires_grad_(True)
    y = _DenseUpsampleFunction.apply(h.cuda(), gg, vv, bb, spec.sub_sr, spec.sub_sr // 2, 1, 0.4)
    assert y.shape == y_ref.shape
    assert torch.allclose(y.cpu(), y_ref.detach(), atol=2e-6, rtol=1e-5)
    (y * w.cuda()).sum().backward()
    for got, ref in zip((gg.grad, vv.grad, bb.grad), grads_ref):
        assert rel_l2(got, ref) < 1e-5
    # without weight norm (remove_weight_norms): v is the plain weight
    weff = O.resolve_weight(sd, "upsampler.1.").cuda().requires_grad_(True)
    y2 = _DenseUpsampleFunction.apply(h.cuda(), None, weff, None, spec.sub_sr, spec.sub_sr // 2, 1, 0.4)
    ref2 = F.leaky_relu(F.conv_transpose1d(F.pad(h, (0, 1), mode="replicate"), weff.detach().cpu(), None,
                                           stride=spec.sub_sr, padding=spec.sub_sr // 2), 0.4)
    assert torch.allclose(y2.cpu(), ref2, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("flip", [False, True])
def test_affine_kernels(flip):
    B, H, W = 3, 16, 37
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, generator=g)
    lst = torch.randn(B, 2, (H - 1) * W, generator=g) * 0.5
    ls, t = lst[:, 0].view(B, H - 1, W), lst[:, 1].view(B, H - 1, W)
    xr = x.clone().requires_grad_(True)
    lr = lst.clone().requires_grad_(True)
    yref = torch.cat((xr[:, :1], xr[:, 1:] * lr[:, 0].view(B, H - 1, W).exp() + lr[:, 1].view(B, H - 1, W)), 1)
    out_ref = yref.flip(1) if flip else yref
    xc, lc = x.cuda(), lst.cuda()
    out = _affine(xc, False, lc, torch.empty_like(xc), flip, False)
    assert torch.allclose(out.cpu(), out_ref.detach(), atol=1e-6, rtol=1e-6)
    # inverse, generated line by line from the (optionally flipped) output
    rec = torch.full_like(xc, float("nan"))
    for j in range(H):
        _affine(out, flip, lc if j else None, rec, False, True, j, 1)
    assert torch.allclose(rec.cpu(), x, atol=2e-6, rtol=1e-5)
    # backward
    dout = torch.randn(B, H, W, generator=g)
    dld = torch.randn(B, generator=g)
    obj = (out_ref * dout).sum() + (lr[:, 0].sum(1) * dld).sum()
    dx_ref, dl_ref = torch.autograd.grad(obj, [xr, lr])
    dx = torch.empty_like(xc)
    dl = torch.empty_like(lc)
    doutc, dldc = dout.cuda(), dld.cuda()
    L.check(L.load().cmwg_waveflow_affine_bwd(xc.data_ptr(), lc.data_ptr(), doutc.data_ptr(), int(flip),
                                              dldc.data_ptr(), dx.data_ptr(), dl.data_ptr(), B, H, W,
                                              L.stream_ptr(xc.device)), "affine_bwd")
    assert torch.allclose(dx.cpu(), dx_ref, atol=1e-6, rtol=1e-5)
    assert torch.allclose(dl.cpu(), dl_ref, atol=2e-6, rtol=1e-5)


def _wn2d_case(n_group, ch, B, H, W, n_mels, seed):
    spec = O.WaveFlowSpec(1, n_group, n_mels)
    sd = O.waveflow_random_state(spec, ch, seed=seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    y = torch.randn(B, n_mels, W, generator=g)
    m = cm.WN2D(n_group, n_mels, **_wkw(ch))
    m.load_state_dict({k[len("WNs.0."):]: v for k, v in sd.items() if k.startswith("WNs.0.")})
    return spec, sd, x, y, m.cuda()


@pytest.mark.parametrize("prec,n_group,ch,H,W", [("fp32", 32, 16, 9, 45), ("fp32", 64, 8, 63, 20), ("fp32", 16, 64, 5, 300),
                                                 ("bf16", 64, 64, 63, 40), ("bf16", 32, 128, 12, 270),
                                                 ("fp16", 64, 64, 20, 33), ("fp16", 64, 64, 63, 40),
                                                 ("fp16", 32, 128, 12, 270)])
def test_wn2d_forward_backward(prec, n_group, ch, H, W):
    B, n_mels = 2, 12
    spec, sd, x, y, m = _wn2d_case(n_group, ch, B, H, W, n_mels, seed=H + W)
    precision.set_precision(prec)
    tol = TOL[prec]
    sdd = to_double({k: v for k, v in sd.items() if k.startswith("WNs.0.")})
    leaf = {k: v.clone().requires_grad_(True) for k, v in sdd.items()}
    xd = x.double().requires_grad_(True)
    yd = y.double().requires_grad_(True)
    ls_ref, t_ref = O.wn2d_forward(leaf, "WNs.0.", xd, yd, spec.h_dilations)
    gen = torch.Generator().manual_seed(5)
    dls = torch.randn(ls_ref.shape, generator=gen)
    dt = torch.randn(t_ref.shape, generator=gen)
    xc = x.cuda().requires_grad_(True)
    yc = y.cuda().requires_grad_(True)
    ls, t = m(xc, yc)
    assert ls.shape == ls_ref.shape
    assert rel_l2(ls, ls_ref) < tol["out"] and rel_l2(t, t_ref) < tol["out"]
    names = [n for n, _ in m.named_parameters()]
    ref = torch.autograd.grad((ls_ref * dls.double()).sum() + (t_ref * dt.double()).sum(),
                              [xd, yd] + [leaf["WNs.0." + n] for n in names])
    ((ls * dls.cuda()).sum() + (t * dt.cuda()).sum()).backward()
    assert rel_l2(xc.grad, ref[0]) < tol["grad_worst"]
    assert rel_l2(yc.grad, ref[1]) < tol["grad_worst"]
    for (n, p), r in zip(m.named_parameters(), ref[2:]):
        if n == "start.weight_v":  # analytically zero (one input channel): rounding noise only
            assert p.grad.abs().max().item() <= 1e-3 * max(1.0, ref[2 + names.index("start.weight_g")].abs().max().item())
            continue
        assert rel_l2(p.grad, r) < tol["grad_worst"], n


@pytest.mark.parametrize("prec,ch", [("fp32", 16), ("bf16", 64), ("fp16", 64)])
def test_wn2d_line_windows_equal_full_forward(prec, ch):
    """The row-recurrent evaluation (one line per call, per-layer state slabs) reproduces the full-image forward bit
    for bit: same tiles, same arithmetic, only the schedule differs."""
    n_group, B, H, W, n_mels = 64, 2, 63, 70, 10
    spec, sd, x, y, m = _wn2d_case(n_group, ch, B, H, W, n_mels, seed=3)
    precision.set_precision(prec)
    lib = L.load()
    img = x.view(B, H, W).cuda().contiguous()
    with torch.no_grad():
        lst_full, st = m._fwd_image(img, H, y.cuda(), save=False, prec=prec)
    cfg = st.cfg
    ws = torch.empty(int(lib.cmwg_wn_workspace_bytes(C.byref(cfg), B, W)), device="cuda", dtype=torch.uint8)
    state = torch.empty(int(lib.cmwg_wn_line_state_bytes(C.byref(cfg), B, W)), device="cuda", dtype=torch.uint8)
    state.fill_(0xFF)  # NaN patterns: lines that are not yet generated must never be read
    lst = torch.full_like(lst_full, float("nan"))
    for h0, nh in [(0, 1), (1, 1), (2, 3)] + [(h, 1) for h in range(5, H)]:
        L.check(lib.cmwg_wn_forward_lines(C.byref(cfg), st.packed.data_ptr(), img.data_ptr(), H * W, st.ycl.data_ptr(),
                                          B, W, h0, nh, ws.data_ptr(), state.data_ptr(), lst.data_ptr(),
                                          L.stream_ptr(img.device)), "wn_forward_lines")
    torch.cuda.synchronize()
    assert torch.equal(lst, lst_full)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_waveflow_against_reference_fixture(tag):
    """fp32 engine against the outputs and gradients of the unmodified reference."""
    fx = load_golden(f"waveflow_tiny_{tag}.pt")
    precision.set_precision("fp32")
    m = cm.WaveFlow(memory_efficient=False, **fx["arch"], **fx["wn_kwargs"])
    m.load_state_dict(fx["state"])
    m = m.cuda().train()
    x, h = fx["x"].cuda(), fx["h"].cuda()
    assert torch.allclose(m._upsample_h(h).cpu(), fx["upsampled"], atol=2e-6, rtol=1e-5)
    z, logdet = m(x, h)
    loss = cm.WaveGlowLoss(fx["sigma"])(z, logdet)
    loss.backward()
    assert rel_l2(z, fx["z"]) < 2e-6
    assert rel_l2(logdet, fx["logdet"]) < 1e-5
    assert abs(loss.item() - fx["loss"].item()) < 1e-5 * abs(fx["loss"].item())
    for n, p in m.named_parameters():
        ref = fx["grads"][n]
        if n.endswith("start.weight_v"):
            assert p.grad.abs().max().item() < 1e-7
            continue
        assert rel_l2(p.grad, ref) < 5e-5, n
    with torch.no_grad():
        xr, ldr = m.reverse(z.detach(), h)
        audio, lds = m.reverse(fx["infer_z"].cuda(), h)
    assert torch.allclose(xr.cpu(), fx["x_roundtrip"], atol=5e-6)
    assert torch.allclose(xr.cpu(), fx["x"], atol=1e-5)
    assert rel_l2(ldr, fx["logdet_reverse"]) < 1e-5
    assert rel_l2(audio, fx["infer_audio"]) < 5e-6
    assert rel_l2(lds, fx["infer_logdet"]) < 1e-5
    assert torch.allclose(m.infer(h, z=fx["infer_z"].cuda()).cpu(), fx["infer_audio"].squeeze(), atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("prec,conv", [("fp32", False), ("fp16", False), ("fp16", True), ("bf16", False), ("bf16", True)])
def test_waveflow_lj_shape_against_oracle(prec, conv):
    """n_group 64 / 64 channels (the LJ config's WN shape), 2 flows, short segment, against the fp64 oracle."""
    precision.set_precision(prec)
    tol = TOL[prec]
    spec = O.WaveFlowSpec(2, 64, 80, conv)
    sd = O.waveflow_random_state(spec, 64, seed=9)
    g = torch.Generator().manual_seed(2)
    B, frames = 2, 8
    x = torch.rand(B, frames * 256, generator=g) * 2 - 1
    h = torch.randn(B, 80, frames, generator=g)
    z_ref, ld_ref, loss_ref, grads_ref = O.waveflow_train_step(to_double(sd), spec, x.double(), h.double(), 0.7)
    m = cm.WaveFlow(2, 64, 80, conv, False, **_wkw(64))
    m.load_state_dict(sd)
    m = m.cuda().train()
    z, logdet = m(x.cuda(), h.cuda())
    loss = cm.WaveGlowLoss(0.7)(z, logdet)
    loss.backward()
    assert rel_l2(z, z_ref) < tol["out"]
    assert rel_l2(logdet, ld_ref) < tol["logdet"]
    worst = 0.0
    for n, p in m.named_parameters():
        if n.endswith("start.weight_v"):
            continue
        worst = max(worst, rel_l2(p.grad, grads_ref[n]))
    assert worst < tol["grad_worst"], worst
    with torch.no_grad():
        xr, ldr = m.reverse(z.detach(), h.cuda())
    assert rel_l2(xr, x) < tol["roundtrip"]
    assert rel_l2(ldr, -ld_ref) < tol["logdet"]
    # synthesis parity on fresh noise
    zs = torch.randn(B, frames * 256, generator=g) * 0.6
    audio_ref, _ = O.waveflow_reverse(to_double(sd), spec, zs.double(), h.double())
    with torch.no_grad():
        audio = m.infer(h.cuda(), z=zs.cuda())
    assert rel_l2(audio, audio_ref.squeeze()) < max(tol["out"], TOL["fp16"]["out"] if prec != "fp32" else 0)


def test_waveflow_synthesis_graph_replay_matches_eager():
    """Second and later calls with the same shapes/weights replay one CUDA graph; results are bit-identical to the
    plain launches, new inputs are honoured, and a weight update invalidates the graph."""
    precision.set_precision("auto")
    spec = O.WaveFlowSpec(2, 64, 80)
    sd = O.waveflow_random_state(spec, 64, seed=4)
    m = cm.WaveFlow(2, 64, 80, False, False, **_wkw(64))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(0)
    h = torch.randn(2, 80, 6, generator=g).cuda()
    z1 = (torch.randn(2, 6 * 256, generator=g) * 0.6).cuda()
    z2 = (torch.randn(2, 6 * 256, generator=g) * 0.6).cuda()
    with torch.no_grad():
        a1, l1 = m.reverse(z1, h)            # plain launches
        before = L.launch_count()
        b1, k1 = m.reverse(z1, h)            # capture + replay
        b2, k2 = m.reverse(z2, h)            # replay with new noise
        assert torch.equal(a1, b1) and torch.equal(l1, k1)
        n_after_capture = L.launch_count()
        c2, _ = m.reverse(z2, h)
        assert L.launch_count() == n_after_capture  # a replay issues no new launches through the library
        assert torch.equal(b2, c2) and not torch.equal(b1, b2)
        assert n_after_capture > before
        with torch.no_grad():
            m.WNs[0].end.weight.mul_(0.5)
        d2, _ = m.reverse(z2, h)             # weights changed: plain launches again, different audio
        assert not torch.equal(d2, c2)
