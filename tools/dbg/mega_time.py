"""Timing + per-role cycle accounting of the forward task kernel at the LJ training shape, per operand precision.
Usage: python tools/dbg/mega_time.py [B] [T]   (CMWG_MEGA_CLK is set here)"""
import os, sys, ctypes as C, torch, numpy as np
os.environ["CMWG_MEGA_CLK"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import _lib
dev = torch.device("cuda", 0)
torch.manual_seed(0)
wn = cm.WN(4, 80, zero_init=False).to(dev)
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 2000)
x = torch.randn(B, 8, T, device=dev); y = torch.randn(B, 80, T, device=dev)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def clk(tag):
    n = 148 * 18 * 16
    buf = (C.c_longlong * n)()
    assert lib.cmwg_mega_clk_read(buf, n) == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(148, 18, 16).astype(np.float64)
    tot = a[:, :, 12]
    print(f"  [{tag}] kernel span (epilogue warps) mean {tot[:, 2:].mean():.0f} clk")
    pr = a[:, 0]; print(f"  producer: wait ring slot {pr[:, 0].mean():.0f}  flag waits {pr[:, 1].mean():.0f}  of {pr[:, 12].mean():.0f}")
    mm = a[0::2, 1]; print(f"  mma (leaders): wait tmem_empty {mm[:, 0].mean():.0f}  wait operands {mm[:, 1].mean():.0f}  of {mm[:, 12].mean():.0f}")
    ep = a[:, 2:]
    for i, nm in enumerate(("G", "R", "S")):
        w, k, s_, c = (ep[:, :, 4 * i + j].mean() for j in range(4))
        print(f"  epilogue {nm}: units/warp {c:.1f}  per unit: wait acc {w / max(c, 1):.0f}  work {k / max(c, 1):.0f}  complete+signal {s_ / max(c, 1):.0f}")


for prec in ("bf16", "fp16"):
    for save in (False, True):
        for _ in range(3):
            wn._cmwg_forward(x, y, save=save, prec=prec)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            wn._cmwg_forward(x, y, save=save, prec=prec)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print(f"{prec} save={save} B={B} T={T}: whole WN forward (start conv + task kernel + end conv) median {ts[5]:.3f} ms  min {ts[0]:.3f}", flush=True)
        clk(f"{prec} save={save}")
