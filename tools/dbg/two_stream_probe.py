"""Does running the WN forward as sub-batch chains (one stream: L2 residency; two streams: wave quantisation) beat one full-batch chain?"""
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
xs = {n: [x[i:i + n].contiguous() for i in range(0, B, n)] for n in (4, 6, 8, 12)}
ys = {n: [y[i:i + n].contiguous() for i in range(0, B, n)] for n in (4, 6, 8, 12)}

def full(save):
    wn._cmwg_forward(x, y, save=save, prec="bf16")

def chunks(n):
    def f(save):
        for a, b in zip(xs[n], ys[n]):
            wn._cmwg_forward(a, b, save=save, prec="bf16")
    return f

def halves(save):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        wn._cmwg_forward(xs[12][0], ys[12][0], save=save, prec="bf16")
    with torch.cuda.stream(s2):
        wn._cmwg_forward(xs[12][1], ys[12][1], save=save, prec="bf16")
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
    print(f"save={save}: full B=24 {timeit(full, save):.3f} ms | 2x12 seq {timeit(chunks(12), save):.3f} | 3x8 seq {timeit(chunks(8), save):.3f} | "
          f"4x6 seq {timeit(chunks(6), save):.3f} | 6x4 seq {timeit(chunks(4), save):.3f} | two streams 2x12 {timeit(halves, save):.3f} ms", flush=True)
