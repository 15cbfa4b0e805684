"""Where does a graphed step diverge from the eager step?  Compares flat gradient buckets, parameters and Adam state."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from constant_memory_waveglow_b200.graphs import GraphedTrainStep
from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets

precision.set_precision("fp32")


def make():
    torch.manual_seed(0)
    m = cm.WaveGlow(4, 8, 2, 2, 256, 80, True, zero_init=False, dilation_channels=64, residual_channels=64,
                    skip_channels=64, depth=2).cuda().train()
    sync = FlowGradSync(flow_buckets(m))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
    return m, sync, opt


loss_fn = cm.WaveGlowLoss(0.7)
ma, sa, oa = make()
mb, sb, ob = make()
mb.load_state_dict(ma.state_dict())
ga = GraphedTrainStep(ma, lambda x, h: loss_fn(*ma(x, h)), oa, sa)
eb = GraphedTrainStep(mb, lambda x, h: loss_fn(*mb(x, h)), ob, sb)
g = torch.Generator(device="cuda").manual_seed(1)
for it in range(3):
    x = torch.rand(3, 4096, device="cuda", generator=g) * 2 - 1
    h = torch.randn(3, 80, 16, device="cuda", generator=g)
    la = ga(x, h).item()
    lb = eb.eager(x, h).item()
    torch.cuda.synchronize()
    print(f"step {it}: loss graph {la:.7f} eager {lb:.7f}")
    worst = []
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        dp = (pa - pb).abs().max().item()
        dg = (pa.grad - pb.grad).abs().max().item() if pa.grad is not None and pb.grad is not None else float("nan")
        sa_, sb_ = oa.state[pa], ob.state[pb]
        dm = (sa_["exp_avg"] - sb_["exp_avg"]).abs().max().item()
        ds = abs(float(sa_["step"]) - float(sb_["step"]))
        worst.append((dp, dg, dm, ds, n))
    worst.sort(reverse=True)
    for w in worst[:6]:
        print("   dparam %.3e dgrad %.3e dexp_avg %.3e dstep %.1f  %s" % w)
    print("   n params with dgrad > 0:", sum(1 for w in worst if w[1] > 0), "of", len(worst))
