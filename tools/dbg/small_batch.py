"""Where a 3-segment training step (the reference's 8-GPU split) spends its time on ONE GPU: eager vs graphed step time and
the per-class device times of the long kernels."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import _lib
from constant_memory_waveglow_b200.graphs import GraphedTrainStep
from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = cm.WaveGlow(memory_efficient=True, zero_init=False, **bench.LJ, **bench.LJ_WN).to(dev).train()
loss_fn = cm.WaveGlowLoss(bench.SIGMA)
sync = FlowGradSync(flow_buckets(model))
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
g = GraphedTrainStep(model, lambda x, h: loss_fn(*model(x, h)), opt, sync)
lib = _lib.load()
for B in (int(a) for a in (sys.argv[1:] or ["3", "6", "12", "24"])):
    x = torch.rand(B, bench.SEGMENT, device=dev) * 2 - 1
    h = torch.randn(B, 80, bench.FRAMES, device=dev)
    def timed(fn, n=10):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    t_e = timed(lambda: g.eager(x, h))
    t_g = timed(lambda: g(x, h))
    lib.cmwg_profile_enable(1)
    g.eager(x, h); torch.cuda.synchronize()
    ms, n = (C.c_double * 8)(), (C.c_longlong * 8)()
    lib.cmwg_profile_collect(ms, n); lib.cmwg_profile_enable(0)
    names = ["gate", "resskip", "dgate", "dx", "dcond", "wgrad", "fwdfused", "bwdfused"]
    cls = {k: (round(ms[i], 3), int(n[i])) for i, k in enumerate(names) if n[i]}
    print(f"B={B}: eager {t_e:.2f} ms/step ({B / t_e * 1e3:.0f} seg/s)  graph {t_g:.2f} ms/step ({B / t_g * 1e3:.0f} seg/s)  "
          f"launches/step {g.launches_per_step}  long kernels: {cls}  sum {sum(v[0] for v in cls.values()):.2f} ms", flush=True)
