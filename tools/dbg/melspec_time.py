"""Device time of the fused log-mel kernel at the training (24 x 16000) and synthesis (1 x 220672) shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from constant_memory_waveglow_b200.condition import MelSpec
m = MelSpec(22050, 1024, 256, f_max=8000, n_mels=80).cuda()
for B, T in ((24, 16000), (1, 220672), (64, 220672)):
    x = torch.rand(B, T, device="cuda") * 2 - 1
    for _ in range(3):
        m(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        m(x)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    byts = B * T * 4 + B * 80 * (T // 256 + 1) * 4
    print(f"B={B} T={T}: {us:.1f} us per call, {byts / us / 1e3:.1f} GB/s algorithmic, {B * (T // 256 + 1)} frames")
