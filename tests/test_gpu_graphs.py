"""graphs.GraphedTrainStep: the whole constant-memory training step replayed as one CUDA graph gives the same parameters,
bit for bit, as the eager step (same kernels in the same stream order), and keeps doing so for new inputs."""
import pytest
import torch

import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from constant_memory_waveglow_b200.graphs import GraphedTrainStep
from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets

pytestmark = pytest.mark.gpu


def _make(seed, wn_ch, depth, flows):
    torch.manual_seed(seed)
    m = cm.WaveGlow(flows, 8, 2, 2, 256, 80, True, zero_init=False, dilation_channels=wn_ch, residual_channels=wn_ch,
                    skip_channels=wn_ch, depth=depth).cuda().train()
    sync = FlowGradSync(flow_buckets(m))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    return m, sync, opt


@pytest.mark.parametrize("prec,wn_ch,depth", [("fp16", 256, 8), ("fp32", 64, 2)])
def test_graphed_step_equals_eager_step(prec, wn_ch, depth):
    old = precision.get_precision()
    precision.set_precision(prec)
    try:
        loss_fn = cm.WaveGlowLoss(0.7)
        ma, sa, oa = _make(0, wn_ch, depth, 4)
        mb, sb, ob = _make(0, wn_ch, depth, 4)
        mb.load_state_dict(ma.state_dict())
        ga = GraphedTrainStep(ma, lambda x, h: loss_fn(*ma(x, h)), oa, sa)
        eb = GraphedTrainStep(mb, lambda x, h: loss_fn(*mb(x, h)), ob, sb)
        g = torch.Generator(device="cuda").manual_seed(1)
        losses = []
        for it in range(4):
            x = torch.rand(3, 4096, device="cuda", generator=g) * 2 - 1
            h = torch.randn(3, 80, 16, device="cuda", generator=g)
            la = ga(x, h).item()            # first call: warm-up (state restored afterwards) + capture + replay; then replays
            lb = eb.eager(x, h).item()
            losses.append((la, lb))
        assert ga.launches_per_step > 50
        for (la, lb) in losses:
            assert la == lb, losses
        for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            assert torch.equal(pa, pb), n
        # a second batch shape: first sight runs eagerly, the second is captured; both stay in step with the eager model
        for it in range(3):
            x = torch.rand(2, 4096, device="cuda", generator=g) * 2 - 1
            h = torch.randn(2, 80, 16, device="cuda", generator=g)
            assert ga(x, h).item() == eb.eager(x, h).item()
        for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            assert torch.equal(pa, pb), n
    finally:
        precision.set_precision(old)


def test_eager_fused_adam_repacks_weights_every_step():
    """Regression: Adam(fused=True) does not bump parameter version counters; the WN forward must still see the updated
    weights.  Two eager models, one stepped by the fused optimizer and one by the plain (version-bumping) implementation,
    stay together."""
    old = precision.get_precision()
    precision.set_precision("fp32")
    try:
        loss_fn = cm.WaveGlowLoss(0.7)
        torch.manual_seed(0)
        kw = dict(zero_init=False, dilation_channels=64, residual_channels=64, skip_channels=64, depth=2)
        ma = cm.WaveGlow(4, 8, 2, 2, 256, 80, True, **kw).cuda().train()
        mb = cm.WaveGlow(4, 8, 2, 2, 256, 80, True, **kw).cuda().train()
        mb.load_state_dict(ma.state_dict())
        oa = torch.optim.Adam(ma.parameters(), lr=1e-3, fused=True)
        ob = torch.optim.Adam(mb.parameters(), lr=1e-3, fused=False, foreach=False)
        g = torch.Generator(device="cuda").manual_seed(3)
        for it in range(4):
            x = torch.rand(2, 4096, device="cuda", generator=g) * 2 - 1
            h = torch.randn(2, 80, 16, device="cuda", generator=g)
            ls = []
            for m, o in ((ma, oa), (mb, ob)):
                o.zero_grad(set_to_none=True)
                loss = loss_fn(*m(x, h))
                loss.backward()
                o.step()
                ls.append(loss.item())
            assert abs(ls[0] - ls[1]) < 1e-5 * abs(ls[1]) + 1e-6, (it, ls)
    finally:
        precision.set_precision(old)
