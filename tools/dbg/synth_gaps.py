"""Synthesis: host time vs device time per call, and where the host time goes."""
import os, sys, time, torch, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench, constant_memory_waveglow_b200 as cm
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = cm.WaveGlow(memory_efficient=True, zero_init=False, **bench.LJ, **bench.LJ_WN).to(dev).eval()
B = 4
hs = torch.randn(B, 80, bench.SYNTH_FRAMES, device=dev); zs = torch.randn(B, bench.SYNTH_FRAMES * 256, device=dev) * 0.6
with torch.no_grad():
    for _ in range(3): model.infer(hs, 0.6, z=zs)
    torch.cuda.synchronize()
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); model.infer(hs, 0.6, z=zs); t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
        print(f"device {a.elapsed_time(b):.2f} ms   host enqueue {(t1 - t0) * 1e3:.2f} ms", flush=True)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): model.infer(hs, 0.6, z=zs)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
