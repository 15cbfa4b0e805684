"""GPU busy / idle accounting of one LJ training step (and one synthesis call) with torch.profiler (CUPTI):
sum of kernel durations, union of busy intervals, idle gaps, per-kernel totals.  Not a benchmark number."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import constant_memory_waveglow_b200 as cm  # noqa: E402
from constant_memory_waveglow_b200 import precision  # noqa: E402
from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (24 if mode == "train" else 4)
precision.set_precision("bf16")
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = cm.WaveGlow(memory_efficient=True, zero_init=False, **bench.LJ, **bench.LJ_WN).to(dev)
if mode == "train":
    model.train()
    loss_fn = cm.WaveGlowLoss(bench.SIGMA)
    sync = FlowGradSync(flow_buckets(model))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.rand(B, bench.SEGMENT, device=dev) * 2 - 1
    h = torch.randn(B, 80, bench.FRAMES, device=dev)

    def step():
        sync.zero_grad()
        z, ld = model(x, h)
        loss_fn(z, ld).backward()
        sync.finish()
        opt.step()
else:
    model.eval()
    hs = torch.randn(B, 80, bench.SYNTH_FRAMES, device=dev)
    zs = torch.randn(B, bench.SYNTH_FRAMES * 256, device=dev) * 0.6

    def step():
        with torch.no_grad():
            model.infer(hs, 0.6, z=zs)

import time  # noqa: E402

for _ in range(3):
    step()
torch.cuda.synchronize()
host = []
for _ in range(5):
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host.append((round((t1 - t0) * 1e3, 2), round((t2 - t0) * 1e3, 2)))
print("host enqueue ms / total ms per step:", host)
if len(sys.argv) > 3 and sys.argv[3] == "cprofile":
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
tot = sum(b - a for a, b, _ in iv)
busy, cur_a, cur_b = 0.0, None, None
gaps = []
for a, b, _ in iv:
    if cur_a is None:
        cur_a, cur_b = a, b
    elif a <= cur_b:
        cur_b = max(cur_b, b)
    else:
        busy += cur_b - cur_a
        gaps.append(a - cur_b)
        cur_a, cur_b = a, b
busy += cur_b - cur_a
span = iv[-1][1] - iv[0][0]
print(f"{mode} B={B}: {len(iv)} device activities, span {span/1e3:.2f} ms, busy {busy/1e3:.2f} ms, "
      f"sum of durations {tot/1e3:.2f} ms, idle {100*(span-busy)/span:.1f} %")
gaps.sort(reverse=True)
print("largest gaps (us):", [round(g, 1) for g in gaps[:12]], " gaps > 5us:", sum(1 for g in gaps if g > 5),
      "total gap us", round(sum(gaps), 1))
agg = defaultdict(lambda: [0, 0.0])
for a, b, n in iv:
    k = n.split("(")[0][-70:]
    agg[k][0] += 1
    agg[k][1] += b - a
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {k:70s} {n:5d} {t/1e3:8.3f} ms {t/n:8.1f} us")
