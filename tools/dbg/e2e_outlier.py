"""Which e2e steps are slow, and does the caching allocator touch the driver in them?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench, constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = cm.WaveGlow(memory_efficient=True, zero_init=False, **bench.LJ, **bench.LJ_WN).to(dev).train()
loss_fn = cm.WaveGlowLoss(bench.SIGMA)
sync = FlowGradSync(flow_buckets(model))
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
B = 24
xh = (torch.rand(B, bench.SEGMENT) * 2 - 1).pin_memory(); hh = torch.randn(B, 80, bench.FRAMES).pin_memory()
xd, hd = xh.to(dev), hh.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step(x, h):
    sync.zero_grad(); z, ld = model(x, h); loss = loss_fn(z, ld); loss.backward(); sync.finish(); opt.step(); return loss
for _ in range(3): step(xd, hd)
torch.cuda.synchronize()
for mode in ("dev", "e2e", "e2e-noflush", "dev"):
    out = []
    for i in range(8):
        if mode != "e2e-noflush": flush.zero_()
        torch.cuda.synchronize()
        s0 = torch.cuda.memory_stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        if mode.startswith("e2e"):
            l = step(xh.to(dev, non_blocking=True), hh.to(dev, non_blocking=True)).item()
        else:
            l = step(xd, hd)
        t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
        s1 = torch.cuda.memory_stats()
        out.append((round(a.elapsed_time(b), 1), round((t1 - t0) * 1e3, 1), s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_device_free"] - s0["num_device_free"]))
    print(mode, out, flush=True)
