"""Single-kernel WN forward (csrc/engine_mega.cuh: every gate / residual / skip GEMM tile of the WN in one dependency-
ordered task list) against the layer-at-a-time pipeline (bit for bit) and the fp64 CPU oracle, over the shapes that
stress its scheduling: one row tile, ragged tiles, fewer tiles than CTA pairs, one layer (no residual tasks), the LJ shape."""
import os

import pytest
import torch

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from oracle import flow_oracle as O
from tests._util import TOL, prefixed, rel_l2, to_double

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore():
    old_p, old_e = precision.get_precision(), os.environ.get("CMWG_MEGA")
    yield
    precision.set_precision(old_p)
    if old_e is None:
        os.environ.pop("CMWG_MEGA", None)
    else:
        os.environ["CMWG_MEGA"] = old_e


def _wn(cin, aux, depth, radix=3, seed=0):
    torch.manual_seed(seed)
    return cm.WN(cin, aux, dilation_channels=256, residual_channels=256, skip_channels=256, depth=depth, radix=radix,
                 zero_init=False).cuda()


def _both(fn):
    """(task kernel, layered pipeline).  The bit-for-bit comparisons keep the `end` conv in its own kernel on both sides
    (CMWG_MEGA_END=0): fused into the skip tiles' epilogue it sums the same products in another order
    (test_fused_end_conv_matches_separate_kernel)."""
    os.environ["CMWG_MEGA_END"] = "0"
    os.environ["CMWG_FOLD0"] = "0"      # ... and the start conv in its own kernel (test_folded_start_conv_matches_separate_kernel)
    os.environ["CMWG_FOLD_END"] = "0"   # ... and the backward chain's dgate tiles on the dskip slab (layered pipeline's operands)
    os.environ["CMWG_MEGA"] = "1"
    try:
        a = fn()
        os.environ["CMWG_MEGA"] = "0"
        b = fn()
    finally:
        os.environ["CMWG_MEGA"] = "1"
        os.environ.pop("CMWG_MEGA_END", None)
        os.environ.pop("CMWG_FOLD0", None)
        os.environ.pop("CMWG_FOLD_END", None)
    return a, b


@pytest.mark.parametrize("cin,aux,depth,radix,B,T", [(4, 80, 8, 3, 2, 300), (2, 80, 8, 3, 3, 2000), (3, 20, 2, 3, 2, 700),
                                                     (4, 80, 8, 3, 24, 2000), (8, 80, 2, 3, 1, 1000), (4, 80, 6, 5, 3, 1500),
                                                     (4, 100, 3, 3, 2, 515)])
def test_folded_start_conv_matches_separate_kernel(cin, aux, depth, radix, B, T):
    """Layer 0 with the start conv folded into its GEMM tiles (W_0 W_start at pack time, the taps of x_a in the padding
    columns of the conditioning slab; the default with fp16 operands) against the separate start conv kernel: the same
    function with other 16-bit roundings (x_a and the folded weights instead of h_0), so the two agree to operand precision;
    both are held to the fp64 oracle.  Ragged tiles, 2 .. 8 input channels, 3 and 5 taps, one and two conditioning k-blocks."""
    wn = _wn(cin, aux, depth, radix=radix, seed=7)
    g = torch.Generator(device="cuda").manual_seed(B * T + cin)
    x = torch.randn(B, 2 * cin, T, device="cuda", generator=g)
    y = torch.randn(B, aux, T, device="cuda", generator=g)

    def run():
        lst, _ = wn._cmwg_forward(x, y, save=False, prec="fp16")
        torch.cuda.synchronize()
        return lst.clone()

    folded = run()
    os.environ["CMWG_FOLD0"] = "0"
    try:
        sep = run()
    finally:
        os.environ.pop("CMWG_FOLD0", None)
    assert torch.isfinite(folded).all()
    assert not torch.equal(folded, sep)       # the fold is on by default at these shapes
    assert rel_l2(folded, sep) < 5e-4, rel_l2(folded, sep)
    assert torch.equal(folded, run())         # deterministic
    sd = to_double({k: v.cpu() for k, v in wn.state_dict().items()})
    log_s, t = O.wn_forward(sd, "", x[:, :cin].double().cpu(), y.double().cpu())
    want = torch.cat([log_s, t], 1)
    assert rel_l2(folded, want) < TOL["fp16"]["out"] and rel_l2(sep, want) < TOL["fp16"]["out"]
    # the saving forward (the recompute of the reversible backward) folds too: the input reconstructed from the output
    # must see the log_s / t the first pass produced
    saved = wn._cmwg_forward(x, y, save=True, prec="fp16")[0]
    assert rel_l2(saved, folded) < 2e-6


@pytest.mark.parametrize("cin,aux,depth,B,T", [(4, 80, 8, 2, 300), (2, 80, 8, 3, 2000), (3, 20, 1, 2, 700), (4, 80, 8, 24, 2000),
                                               (8, 80, 2, 1, 1000)])
@pytest.mark.parametrize("save", [False, True])
def test_fused_end_conv_matches_separate_kernel(cin, aux, depth, B, T, save):
    """`end` 1x1 conv in the epilogue of the skip tiles (default) against the separate kernel reading the fp32 skip slab: the
    same fp32 products, summed in a different order; ragged last tiles, 2 .. 16 output channels, with and without saves."""
    wn = _wn(cin, aux, depth, seed=11)
    g = torch.Generator(device="cuda").manual_seed(B * T)
    x = torch.randn(B, 2 * cin, T, device="cuda", generator=g)
    y = torch.randn(B, aux, T, device="cuda", generator=g)

    def run():
        lst, st = wn._cmwg_forward(x, y, save=save, prec="fp16")
        torch.cuda.synchronize()
        return lst.clone()

    fused = run()
    os.environ["CMWG_MEGA_END"] = "0"
    try:
        sep = run()
    finally:
        os.environ.pop("CMWG_MEGA_END", None)
    assert torch.isfinite(fused).all()
    assert rel_l2(fused, sep) < 2e-6, rel_l2(fused, sep)
    assert torch.equal(fused, run())          # deterministic


SHAPES = [
    # cin, aux, depth, B, T
    (4, 80, 8, 1, 100),     # one row tile: every dependency is the task's own predecessor
    (4, 80, 8, 2, 300),     # two ragged tiles per item, residual tiles directly behind their gate tiles (lag 0)
    (3, 20, 1, 2, 700),     # one layer: gate + skip tasks only
    (4, 80, 2, 3, 1000),
    (2, 40, 5, 5, 2000),    # 40 row tiles < 74 CTA pairs
    (4, 80, 8, 24, 2000),   # the LJ training shape: 192 row tiles
    (4, 80, 8, 2, 27584),   # a 10 s utterance: 108 tiles per item
]


def test_fused_forward_radix5_and_two_streams():
    """Five taps (the halo still reaches one tile at most: 2 * 2^5 = 64 rows), and two streams at once: each stream
    has its own persistent workspace, so concurrent calls must not disturb each other."""
    wn = _wn(4, 80, 6, radix=5, seed=2)
    x = torch.randn(3, 8, 1500, device="cuda")
    y = torch.randn(3, 80, 1500, device="cuda")
    a, b = _both(lambda: wn._cmwg_forward(x, y, save=False, prec="bf16")[0].clone())
    assert torch.equal(a, b)

    wn8 = _wn(4, 80, 8)
    xs = [torch.randn(6, 8, 2000, device="cuda") for _ in range(2)]
    ys = [torch.randn(6, 80, 2000, device="cuda") for _ in range(2)]
    want = [wn8._cmwg_forward(xs[i], ys[i], save=False, prec="bf16")[0].clone() for i in range(2)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [None, None]
    os.environ["CMWG_COOP"] = "1"      # concurrent task kernels: cooperative launches keep each grid whole
    for rep in range(3):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                got[i] = wn8._cmwg_forward(xs[i], ys[i], save=False, prec="bf16")[0]
        torch.cuda.synchronize()
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    os.environ.pop("CMWG_COOP", None)


@pytest.mark.parametrize("cin,aux,depth,B,T", SHAPES)
@pytest.mark.parametrize("prec,save", [("bf16", False), ("bf16", True), ("fp16", False), ("fp16", True)])
def test_fused_forward_equals_layered_pipeline(cin, aux, depth, B, T, prec, save):
    wn = _wn(cin, aux, depth)
    g = torch.Generator(device="cuda").manual_seed(T + B)
    x = torch.randn(B, 2 * cin, T, device="cuda", generator=g)
    y = torch.randn(B, aux, T, device="cuda", generator=g)

    def run():
        lst, st = wn._cmwg_forward(x, y, save=save, prec=prec)
        torch.cuda.synchronize()
        return lst.clone()

    a, b = _both(run)       # (the saved activations are compared through the gradients they produce, below)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
    c = run()                            # default arrangement (`end` conv in the skip tiles' epilogue, start conv folded
    assert torch.equal(c, run())         # into layer 0); run to run: same bits (fixed accumulation order, no atomics on data)
    folded = prec == "fp16" and depth >= 2
    assert rel_l2(c, a) < (5e-4 if folded else 2e-6)


@pytest.mark.parametrize("cin,aux,depth,B,T", [(4, 80, 8, 2, 300), (3, 20, 1, 2, 700), (4, 80, 4, 2, 1000)])
def test_fused_forward_against_oracle(cin, aux, depth, B, T):
    """log_s / t of the WN against the fp64 restatement of model/waveglow.py:98-105, operand-precision tolerance."""
    wn = _wn(cin, aux, depth, seed=3)
    torch.manual_seed(5)
    x = torch.randn(B, 2 * cin, T)
    y = torch.randn(B, aux, T)
    sd = to_double({k: v.cpu() for k, v in wn.state_dict().items()})
    log_s, t = O.wn_forward(sd, "", x[:, :cin].double(), y.double())
    want = torch.cat([log_s, t], 1)
    for prec in ("bf16", "fp16"):
        lst, _ = wn._cmwg_forward(x.cuda(), y.cuda(), save=False, prec=prec)
        assert rel_l2(lst, want) < TOL[prec]["out"], prec


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_training_step_gradients_equal_layered_pipeline(prec):
    """Reversible backward through the activations the fused forward saved: every gradient equals the layered
    pipeline's bit for bit (the backward kernels are the same; their inputs must be)."""
    torch.manual_seed(0)
    blk = cm.AffineCouplingBlock(cm.WN, True, in_channels=4, aux_channels=80, zero_init=False, dilation_channels=256,
                                 residual_channels=256, skip_channels=256, depth=8).cuda()
    x0 = torch.rand(3, 8, 2000, device="cuda") * 2 - 1
    y = torch.randn(3, 80, 2000, device="cuda")
    precision.set_precision(prec)

    def run():
        for p in blk.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        xin = x * 1.0
        z, log_s = blk(xin, y)
        (z.square().mean() + log_s.mean()).backward()
        torch.cuda.synchronize()
        return [x.grad.clone()] + [p.grad.clone() for p in blk.parameters()]

    os.environ["CMWG_WGRAD_BATCH"] = "0"         # same split-K plan as the layered pipeline: bit for bit
    try:
        ga, gb = _both(run)
    finally:
        os.environ.pop("CMWG_WGRAD_BATCH", None)
    assert len(ga) == len(gb) and all(torch.equal(a, b) for a, b in zip(ga, gb))
    assert all(torch.isfinite(a).all() for a in ga)
    # default: the weight-gradient GEMMs of all layers in one launch, full-K tiles instead of split-K partials --
    # the same products summed in another order
    gc = run()
    # ... and the `end` conv sits in the skip tiles' epilogue (its fp32 products summed in another order: log_s / t move by
    # ~1e-7, which flips a few 16-bit roundings of the gradient slabs downstream), so every gradient agrees to well below
    # the operand precision rather than bit for bit
    # ... and with fp16 operands layer 0 runs without the start conv (x_a and W_0 W_start are rounded to 16 bits instead of
    # h_0; the weight gradient of its dilated conv goes through the fold as well)
    for a, c in zip(ga, gc):
        assert rel_l2(c, a) < (6e-4 if prec == "fp16" else 4e-3), rel_l2(c, a)
    gd = run()
    assert all(torch.equal(c, d) for c, d in zip(gc, gd))      # the default arrangement is deterministic too


@pytest.mark.parametrize("end_scale", [1.0, 1e-5])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_folded_end_conv_backward_matches_dskip_slab(prec, end_scale):
    """Backward chain with the `end` conv folded into the dgate tiles ((W_end W_skip)^T against a one-k-block slab of
    S * d(log_s, t); the default) against the tiles that read the 256-channel dskip slab: the same function with other 16-bit
    roundings.  A nearly-zero `end` weight (zero_init models a few steps into training) makes S * d(log_s, t) itself the
    binding magnitude for the power-of-two gradient scale."""
    torch.manual_seed(1)
    blk = cm.AffineCouplingBlock(cm.WN, True, in_channels=4, aux_channels=80, zero_init=False, dilation_channels=256,
                                 residual_channels=256, skip_channels=256, depth=8).cuda()
    with torch.no_grad():
        blk.F.end.weight.mul_(end_scale)
    x0 = torch.rand(3, 8, 1800, device="cuda") * 2 - 1
    y = torch.randn(3, 80, 1800, device="cuda")
    precision.set_precision(prec)

    def run():
        for p in blk.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        z, log_s = blk(x * 1.0, y)
        (z.square().mean() + log_s.mean()).backward()
        torch.cuda.synchronize()
        return [x.grad.clone()] + [p.grad.clone() for p in blk.parameters()]

    folded = run()
    os.environ["CMWG_FOLD_END"] = "0"
    try:
        slab = run()
    finally:
        os.environ.pop("CMWG_FOLD_END", None)
    assert all(torch.isfinite(a).all() for a in folded)
    num = sum((a.double() - b.double()).square().sum() for a, b in zip(folded, slab)).sqrt()
    den = sum(b.double().square().sum() for b in slab).sqrt()
    assert num / den < (6e-4 if prec == "fp16" else 4e-3), float(num / den)
    assert any(not torch.equal(a, b) for a, b in zip(folded, slab))      # the fold is on by default
    again = run()
    assert all(torch.equal(a, b) for a, b in zip(folded, again))          # deterministic


def test_activation_storing_model_shares_the_conditioning_slab():
    """memory_efficient=False runs every flow's forward (with saves) before the first backward.  The taps of x_a that the folded
    start conv reads ride in padding columns of the conditioning slab ALL flows share, so by the time flow 0's backward runs
    they hold the last flow's taps -- the backward writes them again (csrc/wn_pipeline.cu).  Same weights, same inputs: the
    activation-storing model and the constant-memory model must produce the same gradients."""
    precision.set_precision("fp16")
    wn_kw = dict(dilation_channels=256, residual_channels=256, skip_channels=256, depth=3)
    torch.manual_seed(4)
    stored = cm.WaveGlow(3, 8, 2, 2, 256, 80, False, zero_init=False, **wn_kw).cuda().train()
    const = cm.WaveGlow(3, 8, 2, 2, 256, 80, True, zero_init=False, **wn_kw).cuda().train()
    const.load_state_dict(stored.state_dict())
    x = torch.rand(2, 8192, device="cuda") * 2 - 1
    h = torch.randn(2, 80, 32, device="cuda")
    grads = []
    for m in (stored, const):
        z, logdet = m(x.clone(), h.clone())
        cm.WaveGlowLoss(0.7)(z, logdet).backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters()})
    for n in grads[0]:
        assert torch.isfinite(grads[0][n]).all(), n
        assert rel_l2(grads[0][n], grads[1][n]) < 1e-4, (n, rel_l2(grads[0][n], grads[1][n]))
