"""Does running the WN forward as two half-batch chains on two streams beat one full-batch chain (wave quantisation)?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
precision.set_precision("bf16")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
wn = cm.WN(4, 80, zero_init=False).to(dev)
B, T = 24, 2000
x = torch.randn(B, 8, T, device=dev)
y = torch.randn(B, 80, T, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def full(save):
    wn._cmwg_forward(x, y, save=save, prec="bf16")

def halves(save):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        wn._cmwg_forward(x[:12], y[:12], save=save, prec="bf16")
    with torch.cuda.stream(s2):
        wn._cmwg_forward(x[12:], y[12:], save=save, prec="bf16")
    cur.wait_stream(s1); cur.wait_stream(s2)

def timeit(fn, save, n=20):
    for _ in range(3): fn(save)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn(save)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for save in (False, True):
    print(f"save={save}: full B=24 {timeit(full, save):.3f} ms   two streams 2 x B=12 {timeit(halves, save):.3f} ms")
