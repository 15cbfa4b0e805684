"""Is the K-major tcgen05 path itself slower than the MN-major one, or is it per-tile overhead?  Plain GEMMs of the
same FLOPs through the engine's self test: K-major with short K (many tiles) vs long K (few tiles) vs MN-major long K."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from constant_memory_waveglow_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (M, N, K) in ((256 * 74 * 8, 256, 896), (256 * 74 * 2, 256, 3584), (256 * 74, 256, 7168), (256 * 74, 256, 16384)):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); b = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    ms = t(lambda: ops.selftest_tc_gemm(a, b, M, N, K, 0))
    print(f"K-major  M={M:7d} N={N} K={K:6d}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    del a, b
for (M, N, K) in ((256 * 74, 256, 16384), (512, 256 * 37, 16384)):
    a = (torch.randn(K, M, device="cuda") * 0.5).bfloat16(); b = (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
    ms = t(lambda: ops.selftest_tc_gemm(a, b, M, N, K, 1))
    print(f"MN-major M={M:7d} N={N} K={K:6d}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    del a, b
