"""One LJ training step (or one synthesis call) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Usage: python tools/profile_step.py [train|trainopt|synth] [precision] [batch]
(trainopt: with the fused Adam step, so the weight packs of all flows are rebuilt inside the profiled step, as in bench.py)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm  # noqa: E402
from constant_memory_waveglow_b200 import precision  # noqa: E402
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 24
precision.set_precision(prec)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = cm.WaveGlow(memory_efficient=True, zero_init=False, **bench.LJ, **bench.LJ_WN).to(dev)
loss_fn = cm.WaveGlowLoss(bench.SIGMA)
if mode in ("train", "trainopt"):
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True) if mode == "trainopt" else None
    x = torch.rand(B, bench.SEGMENT, device=dev) * 2 - 1
    h = torch.randn(B, 80, bench.FRAMES, device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        z, ld = model(x, h)
        loss_fn(z, ld).backward()
        if opt is not None:
            opt.step()
else:
    model.eval()
    hs = torch.randn(B, 80, bench.SYNTH_FRAMES, device=dev)
    zs = torch.randn(B, bench.SYNTH_FRAMES * 256, device=dev) * 0.6

    def step():
        with torch.no_grad():
            model.infer(hs, 0.6, z=zs)

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one", mode, "step")
