"""WN transform (forward + backward) and the affine coupling block through the CUDA path against
the fp64 CPU oracle, for the exact fp32 engine and the tcgen05 engine, plus the golden fixtures the
unmodified reference produced."""
import pytest
import torch

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from oracle import flow_oracle as O
from tests._util import TOL, grad_errors, load_golden, prefixed, rel_l2, to_double

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    old = precision.get_precision()
    yield
    precision.set_precision(old)


def make_block(cin, aux, ch, depth, efficient, seed=0, bias=False, radix=3):
    torch.manual_seed(seed)
    blk = cm.AffineCouplingBlock(cm.WN, efficient, in_channels=cin, aux_channels=aux, zero_init=False,
                                 dilation_channels=ch, residual_channels=ch, skip_channels=ch, depth=depth,
                                 bias=bias, radix=radix)
    if bias:
        with torch.no_grad():
            blk.F.end.bias.normal_(std=0.05)
    return blk


CASES = [
    # cin, aux, ch, depth, B, T
    (8, 20, 32, 2, 2, 300),
    (4, 12, 32, 3, 2, 211),
    (8, 40, 128, 4, 2, 4000),   # the reference test grid's largest block (tests/test_fwd_bwd.py:82-88)
    (16, 20, 128, 1, 2, 1000),
    (4, 80, 64, 3, 1, 700),
]


def run_case(prec, cin, aux, ch, depth, B, T, direction, bias=False, radix=3):
    precision.set_precision(prec)
    blk = make_block(cin, aux, ch, depth, True, seed=cin + depth, bias=bias, radix=radix)
    sd = prefixed(blk.state_dict(), "")
    sd64 = to_double({k: v.clone() for k, v in sd.items()})
    g = torch.Generator().manual_seed(100 + T)
    x = torch.rand(B, 2 * cin, T, generator=g, dtype=torch.float64) * 2 - 1
    y = torch.randn(B, aux, T, generator=g, dtype=torch.float64)
    dz = torch.randn(B, 2 * cin, T, generator=g, dtype=torch.float64) / (B * T)
    dls = torch.full((B, cin, T), -1.0 / (B * T), dtype=torch.float64)
    rev = direction == "reverse"
    out_ref, ls_ref, dx_ref, dp_ref, dy_ref = O.coupling_grads(sd64, "F.", x, y, dz, dls, reverse=rev, need_dy=True)

    blk = blk.cuda()
    xg = x.float().cuda().requires_grad_(True)
    yg = y.float().cuda().requires_grad_(True)
    xin = xg.clone()
    out, ls = (blk.reverse(xin, yg) if rev else blk(xin, yg))
    assert xin.untyped_storage().size() == 0            # input consumed (efficient_modules.py:74)
    tol = TOL[prec]
    assert rel_l2(out, out_ref) < tol["out"], ("out", rel_l2(out, out_ref))
    assert rel_l2(ls, ls_ref) < max(tol["out"], 1e-5), ("log_s", rel_l2(ls, ls_ref))
    obj = (out * dz.float().cuda()).sum() + (ls * dls.float().cuda()).sum()
    obj.backward()
    assert xin.untyped_storage().size() > 0              # input re-materialised in place
    assert rel_l2(xin, x) < tol["roundtrip"], ("restore", rel_l2(xin, x))
    assert rel_l2(xg.grad, dx_ref) < tol["grad_worst"], ("dx", rel_l2(xg.grad, dx_ref))
    assert rel_l2(yg.grad, dy_ref) < tol["grad_worst"], ("dy", rel_l2(yg.grad, dy_ref))
    agg, worst, name = grad_errors([(n, p.grad) for n, p in blk.named_parameters()], dp_ref)
    assert agg < tol["grad"], ("aggregate", agg)
    assert worst < tol["grad_worst"], (name, worst)
    return worst


@pytest.mark.parametrize("direction", ["forward", "reverse"])
@pytest.mark.parametrize("case", CASES)
def test_coupling_wn_fp32_engine(case, direction):
    run_case("fp32", *case, direction)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("direction", ["forward", "reverse"])
@pytest.mark.parametrize("case", [c for c in CASES if c[2] % 64 == 0])
def test_coupling_wn_tensor_core(case, direction, prec):
    run_case(prec, *case, direction)


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3e4])
def test_fp16_gradient_scale_is_magnitude_independent(scale):
    """The fp16-operand backward runs on S * cotangent with S chosen on the device (grad_scale_kernel): cotangents six orders
    of magnitude below / four above O(1) -- far outside what fp16 could hold unscaled -- give the same relative accuracy."""
    precision.set_precision("fp16")
    cin, aux, ch, depth, B, T = 4, 80, 256, 3, 2, 500
    blk = make_block(cin, aux, ch, depth, True, seed=5)
    sd64 = to_double({k: v.clone() for k, v in blk.state_dict().items()})
    g = torch.Generator().manual_seed(9)
    x = torch.rand(B, 2 * cin, T, generator=g, dtype=torch.float64) * 2 - 1
    y = torch.randn(B, aux, T, generator=g, dtype=torch.float64)
    dz = torch.randn(B, 2 * cin, T, generator=g, dtype=torch.float64) * scale
    dls = torch.randn(B, cin, T, generator=g, dtype=torch.float64) * scale
    _, _, dx_ref, dp_ref, dy_ref = O.coupling_grads(sd64, "F.", x, y, dz, dls, reverse=False, need_dy=True)
    blk = blk.cuda()
    xg = x.float().cuda().requires_grad_(True)
    yg = y.float().cuda().requires_grad_(True)
    out, ls = blk(xg.clone(), yg)
    ((out * dz.float().cuda()).sum() + (ls * dls.float().cuda()).sum()).backward()
    tol = TOL["fp16"]
    agg, worst, name = grad_errors([(n, p.grad) for n, p in blk.named_parameters()], dp_ref)
    assert agg < tol["grad"] and worst < tol["grad_worst"], (agg, name, worst)
    assert rel_l2(xg.grad, dx_ref) < tol["grad_worst"] and rel_l2(yg.grad, dy_ref) < tol["grad_worst"]


# 256 channels: the start / end conv backward take their one-pass fast paths (start_bwd256_kernel for in_channels
# <= 8, end_bwd_dw256_kernel for in_channels <= 4); row counts that are not multiples of 32 make the 32-row warp
# groups straddle batch items and leave ragged tails
FAST256 = [(4, 80, 256, 2, 2, 333), (3, 80, 256, 1, 3, 200), (2, 16, 256, 1, 2, 97), (7, 24, 256, 1, 2, 150),
           (2, 80, 256, 1, 1, 2000),
           (8, 200, 256, 2, 2, 300)]   # WSRGlow-like: 8 input channels (16 outputs of `end`), four conditioning k-blocks -- the
                                       # widest shapes the folded start / end convs take (csrc/wn_pipeline.cu: fold0, fe_full)


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("case", FAST256)
def test_coupling_wn_256_channel_fast_paths(case, prec):
    run_case(prec, *case, "forward")


def test_wn_256_channel_fast_paths_bias():
    run_case("fp32", 2, 12, 256, 1, 2, 130, "forward", bias=True)
    run_case("bf16", 4, 12, 256, 1, 2, 130, "reverse", bias=True)
    run_case("fp16", 4, 12, 256, 1, 2, 130, "reverse", bias=True)


def test_wn_bias_and_radix5_fp32():
    run_case("fp32", 4, 12, 32, 2, 2, 150, "forward", bias=True, radix=5)
    run_case("fp32", 4, 12, 64, 2, 1, 200, "reverse", bias=True)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_wn_bias_tensor_core(prec):
    run_case(prec, 4, 12, 64, 2, 2, 300, "forward", bias=True)


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
def test_wn_forward_lj_layer_shapes(prec):
    """One flow of the LJ configuration (256 channels, 8 layers, 80 mel), forward only."""
    precision.set_precision(prec)
    torch.manual_seed(3)
    wn = cm.WN(4, 80, zero_init=False)
    sd64 = to_double(prefixed({k: v.clone() for k, v in wn.state_dict().items()}, "F."))
    B, T = 2, 1000
    x = torch.rand(B, 4, T, dtype=torch.float64) * 2 - 1
    y = torch.randn(B, 80, T, dtype=torch.float64)
    ls_ref, t_ref = O.wn_forward(sd64, "F.", x, y)
    wn = wn.cuda()
    with torch.no_grad():
        ls, t = wn(x.float().cuda(), y.float().cuda())
        ls2, t2 = wn(x.float().cuda(), y.float().cuda())
    assert torch.equal(ls, ls2) and torch.equal(t, t2)       # bitwise deterministic
    tol = TOL[prec]["out"]
    assert rel_l2(ls, ls_ref) < tol and rel_l2(t, t_ref) < tol, (rel_l2(ls, ls_ref), rel_l2(t, t_ref))


@pytest.mark.parametrize("name", ["a", "b"])
def test_coupling_against_reference_golden(name):
    """Fixtures produced by the unmodified reference's memory-efficient path (fp32 CPU)."""
    precision.set_precision("fp32")
    fx = load_golden(f"coupling_{name}.pt")
    for direction in ("forward", "reverse"):
        blk = cm.AffineCouplingBlock(cm.WN, True, **fx["kwargs"])
        blk.load_state_dict(fx["state"])
        blk = blk.cuda()
        ref = fx[direction]
        x = fx["x"].cuda().requires_grad_(True)
        y = fx["y"].cuda().requires_grad_(True)
        B = x.shape[0]
        out, ls = (blk(x.clone(), y) if direction == "forward" else blk.reverse(x.clone(), y))
        assert rel_l2(out, ref["out"]) < 2e-6 and rel_l2(ls, ref["log_s"]) < 1e-5
        loss = cm.WaveGlowLoss(1.0)(out.reshape(B, -1), ls.sum((1, 2)))
        loss.backward()
        assert abs(loss.item() - ref["loss"].item()) < 1e-5 * abs(ref["loss"].item())
        assert rel_l2(x.grad, ref["dx"]) < 2e-5 and rel_l2(y.grad, ref["dy"]) < 2e-5
        for n, p in blk.named_parameters():
            assert rel_l2(p.grad, ref["dparams"][n]) < 5e-5, n


def test_generic_transform_falls_back_to_module_call():
    """AffineCouplingBlock is generic over transform_type (melglow.py:197): a non-WN transform is
    called as a module and differentiated with autograd, with the same free/restore contract."""
    class Tiny(torch.nn.Module):
        def __init__(self, c, aux):
            super().__init__()
            self.a = torch.nn.Conv1d(c + aux, 2 * c, 3, padding=1)

        def forward(self, x, y):
            return (0.3 * torch.tanh(self.a(torch.cat((x, y), 1)))).chunk(2, 1)

    torch.manual_seed(0)
    eff = cm.AffineCouplingBlock(Tiny, True, c=3, aux=5).cuda()
    nai = cm.AffineCouplingBlock(Tiny, False, c=3, aux=5).cuda()
    nai.load_state_dict(eff.state_dict())
    x = torch.randn(2, 6, 50).cuda()
    y = torch.randn(2, 5, 50).cuda()
    res = []
    for blk in (eff, nai):
        blk.zero_grad()
        xin = x.clone().requires_grad_(True)
        xc = xin.clone()
        z, ls = blk(xc, y)
        ((z ** 2).sum() + ls.sum()).backward()
        res.append((z.detach(), xin.grad, [p.grad.clone() for p in blk.parameters()]))
        if blk is eff:
            assert torch.allclose(xc, x, atol=1e-6)
    assert torch.allclose(res[0][0], res[1][0], atol=1e-6)
    assert torch.allclose(res[0][1], res[1][1], atol=1e-5)
    for a, b in zip(res[0][2], res[1][2]):
        assert torch.allclose(a, b, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("cin,aux,ch,depth,radix,height", [(4, 80, 256, 8, 3, 0), (3, 20, 64, 2, 5, 0), (8, 100, 128, 3, 3, 0),
                                                          (1, 80, 64, 3, 3, 8)])
def test_vector_operand_pack_equals_scalar_pack(cin, aux, ch, depth, radix, height, prec):
    """csrc/wn_kernels.cuh: pack_operands16_kernel (8 / 16 elements per thread, coalesced reads for the transposed
    matrices) writes the same bytes as the one-element-per-thread kernel; 1-D and 2-D (WaveFlow) weight layouts."""
    import os
    torch.manual_seed(cin * 100 + aux)
    if height:
        from constant_memory_waveglow_b200.waveflow import WN2D
        wn = WN2D(height, aux, dilation_channels=ch, residual_channels=ch, skip_channels=ch, zero_init=False).cuda()
    else:
        wn = cm.WN(cin, aux, dilation_channels=ch, residual_channels=ch, skip_channels=ch, depth=depth, radix=radix,
                   zero_init=False).cuda()
    cm.invalidate_packs()
    a = wn._prepare(prec, torch.device("cuda", 0), height)[1].clone()
    os.environ["CMWG_PACK_SCALAR"] = "1"
    try:
        cm.invalidate_packs()
        b = wn._prepare(prec, torch.device("cuda", 0), height)[1].clone()
    finally:
        os.environ.pop("CMWG_PACK_SCALAR", None)
        cm.invalidate_packs()
    assert a.numel() == b.numel() and torch.equal(a, b)
