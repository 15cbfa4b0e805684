"""HBM-bound primitives (1x1 conv, coupling, squeeze, loss, upsampler) through the C ABI against the
fp64 CPU oracle.  fp32 kernels: tolerances are a few ulp of fp32."""
import pytest
import torch
import torch.nn.functional as F

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import ops
from oracle import flow_oracle as O
from tests._util import rel_l2, max_abs

pytestmark = pytest.mark.gpu


def _orth(c, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.linalg.qr(torch.randn(c, c, generator=g, dtype=torch.float64))[0]
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q + 0.05 * torch.randn(c, c, generator=g, dtype=torch.float64)


@pytest.mark.parametrize("c", [2, 4, 6, 8, 12, 16, 24])
@pytest.mark.parametrize("B,T", [(1, 2000), (3, 257), (2, 4)])
def test_conv1x1_apply_and_inverse(c, B, T):
    w = _orth(c, c)
    x = torch.rand(B, c, T, dtype=torch.float64) * 2 - 1
    z_ref, ld_ref = O.conv1x1_forward(w.unsqueeze(-1), x)
    xr_ref, _ = O.conv1x1_reverse(w.unsqueeze(-1), x)
    wg, xg = w.float().cuda(), x.float().cuda()
    winv, logdet = ops.small_inverse_logdet(wg)
    assert rel_l2(winv, w.inverse()) < 1e-6
    assert abs(logdet.item() - w.logdet().item()) < 1e-6 * max(1, abs(w.logdet().item())) + 1e-6
    assert rel_l2(ops.conv1x1_apply(wg, xg), z_ref) < 1e-6
    assert rel_l2(ops.conv1x1_apply(winv, xg), xr_ref) < 2e-6
    assert rel_l2(ops.conv1x1_apply(wg, xg, transpose=True), F.conv1d(x, w.t().unsqueeze(-1))) < 1e-6
    # channel-slice input (batch stride larger than C*T) is consumed without a copy
    wide = torch.rand(B, c + 2, T).cuda()
    sl = wide[:, :c]
    assert rel_l2(ops.conv1x1_apply(wg, sl), F.conv1d(sl.double().cpu(), w.unsqueeze(-1))) < 1e-6


def test_logdet_negative_determinant_is_nan():
    w = torch.eye(4).flip(0)[[0, 1, 3, 2]].cuda()  # a permutation with det = -1 ... or +1; force sign below
    w = torch.diag(torch.tensor([1.0, 1.0, 1.0, -1.0])).cuda()
    _, ld = ops.small_inverse_logdet(w)
    assert torch.isnan(ld)


@pytest.mark.parametrize("c", [2, 8, 16])
@pytest.mark.parametrize("inverse", [False, True])
def test_conv1x1_weight_gradient(c, inverse):
    B, T = 3, 1300
    w = _orth(c, 10 + c)
    x = torch.rand(B, c, T, dtype=torch.float64) * 2 - 1
    dz = torch.randn(B, c, T, dtype=torch.float64)
    dl = torch.tensor(-0.37, dtype=torch.float64)
    fn = O.invconv1x1_backward if inverse else O.conv1x1_backward
    dx_ref, dw_ref = fn(w.unsqueeze(-1), x, dz, dl)
    wg = w.float().cuda()
    winv, _ = ops.small_inverse_logdet(wg)
    dm = ops.conv1x1_wgrad(dz.float().cuda(), x.float().cuda())
    dw = ops.conv1x1_dw_finalize(dm, winv, dl.float().cuda(), T, inverse)
    assert rel_l2(dw, dw_ref.squeeze(-1)) < 5e-6
    dx = ops.conv1x1_apply(winv if inverse else wg, dz.float().cuda(), transpose=True)
    assert rel_l2(dx, dx_ref) < 2e-6


@pytest.mark.parametrize("cin", [1, 2, 4, 8])
@pytest.mark.parametrize("B,T", [(2, 2000), (1, 37)])
def test_coupling_apply_and_backward(cin, B, T):
    g = torch.Generator().manual_seed(cin * 7 + T)
    x = torch.rand(B, 2 * cin, T, generator=g, dtype=torch.float64) * 2 - 1
    lst = torch.randn(B, 2 * cin, T, generator=g, dtype=torch.float64) * 0.5
    ls, t = lst.chunk(2, 1)
    xa, xb = x.chunk(2, 1)
    z_ref = torch.cat((xa, xb * ls.exp() + t), 1)
    xi_ref = torch.cat((xa, (xb - t) / ls.exp()), 1)
    xg, lg = x.float().cuda(), lst.float().cuda()
    z, none = ops.coupling_apply(xg, lg, False)
    assert none is None and rel_l2(z, z_ref) < 1e-6
    xi, neg = ops.coupling_apply(xg, lg, True)
    assert rel_l2(xi, xi_ref) < 1e-6 and torch.equal(neg, -lg[:, :cin])
    # backward of the forward direction: restore x, cotangent of (log_s, t), d(xb)
    dz = torch.randn(B, 2 * cin, T, generator=g, dtype=torch.float64)
    dls = torch.randn(B, cin, T, generator=g, dtype=torch.float64)
    restored = torch.empty_like(z)
    dlst, din = ops.coupling_bwd(z, lg, dz.float().cuda(), dls.float().cuda(), restored, False)
    assert max_abs(restored, x) < 2e-6
    dza, dzb = dz.chunk(2, 1)
    assert rel_l2(dlst, torch.cat((dzb * xb * ls.exp() + dls, dzb), 1)) < 2e-6
    assert rel_l2(din, torch.cat((dza, dzb * ls.exp()), 1)) < 1e-6
    # backward of the inverse direction
    restored2 = torch.empty_like(z)
    dlst2, din2 = ops.coupling_bwd(xi, lg, dz.float().cuda(), dls.float().cuda(), restored2, True)
    assert max_abs(restored2, x) < 2e-6
    xbi = (xb - t) / ls.exp()
    assert rel_l2(dlst2, torch.cat((-dzb * xbi - dls, -dzb / ls.exp()), 1)) < 2e-6
    assert rel_l2(din2, torch.cat((dza, dzb / ls.exp()), 1)) < 1e-6


def test_squeeze_roundtrip_and_layout():
    x = torch.randn(3, 16000).cuda()
    s = ops.squeeze(x, 8)
    assert torch.equal(s, x.view(3, -1, 8).transpose(1, 2).contiguous())
    assert torch.equal(ops.squeeze(s, 8, inverse=True), x)


@pytest.mark.parametrize("B,T,mean", [(2, 16000, True), (5, 333, False)])
def test_nll_loss_and_gradient(B, T, mean):
    z = torch.randn(B, T, dtype=torch.float64)
    ld = torch.randn(B, dtype=torch.float64) * 10
    zz = z.clone().requires_grad_(True)
    ll = ld.clone().requires_grad_(True)
    ref = O.waveglow_loss(zz, ll, 0.7, mean)
    gz, gl = torch.autograd.grad(ref, [zz, ll])
    zc = z.float().cuda().requires_grad_(True)
    lc = ld.float().cuda().requires_grad_(True)
    loss = cm.WaveGlowLoss(0.7, mean)(zc, lc)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
    assert rel_l2(zc.grad, gz) < 1e-6 and rel_l2(lc.grad, gl) < 1e-6


@pytest.mark.parametrize("C,F_,K,stride,pad", [(80, 63, 65, 32, 16), (8, 8, 65, 32, 16), (5, 40, 3, 1, 1), (3, 7, 9, 4, 2),
                                                 (2, 13000, 3, 1, 1)])   # a frame row longer than the 48 KB staging buffer
def test_upsampler_forward_backward(C, F_, K, stride, pad):
    g0 = torch.Generator().manual_seed(C + K)
    B = 2
    h = torch.randn(B, C, F_, generator=g0, dtype=torch.float64)
    v = torch.randn(C, 1, K, generator=g0, dtype=torch.float64) * 0.2
    gg = torch.rand(C, 1, 1, generator=g0, dtype=torch.float64) + 0.5
    bias = torch.randn(C, generator=g0, dtype=torch.float64)
    vv, g2, b2 = (t.clone().requires_grad_(True) for t in (v, gg, bias))
    y_ref = F.conv_transpose1d(h, O.weight_norm_weight(g2, vv), b2, stride=stride, padding=pad, groups=C)
    Tv = y_ref.shape[-1] - 1 if y_ref.shape[-1] > 1 else 1
    dy = torch.randn(B, C, Tv, generator=g0, dtype=torch.float64)
    (y_ref[..., :Tv] * dy).sum().backward()
    y = ops.upsample_fwd(h.float().cuda(), gg.float().cuda(), v.float().cuda(), bias.float().cuda(), stride, pad)
    assert y.shape == y_ref.shape and rel_l2(y, y_ref) < 1e-6
    dg, dv, db = ops.upsample_bwd(h.float().cuda(), gg.float().cuda(), v.float().cuda(), dy.float().cuda(), stride, pad, True)
    assert rel_l2(dv, vv.grad) < 5e-6 and rel_l2(dg, g2.grad) < 5e-6 and rel_l2(db, b2.grad) < 5e-6
