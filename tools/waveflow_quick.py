"""WaveFlow LJ config (configs/waveflow_LJ_speech.json of the reference): time one training step (forward + loss +
backward) and one synthesis call on cuda:0.  Usage: python tools/waveflow_quick.py [train_batch] [synth_batch] [prec]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm  # noqa: E402
from constant_memory_waveglow_b200 import _lib, precision  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
SB = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "auto"
precision.set_precision(prec)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
m = cm.WaveFlow(flows=8, n_group=64, n_mels=80, use_conv1x1=False, memory_efficient=False, dilation_channels=64,
                residual_channels=64, skip_channels=64, bias=False, zero_init=False).to(dev)
loss_fn = cm.WaveGlowLoss(0.7)
x = torch.rand(B, 16000, device=dev) * 2 - 1
h = torch.randn(B, 80, 63, device=dev)


def step():
    m.zero_grad(set_to_none=True)
    z, ld = m(x, h)
    loss = loss_fn(z, ld)
    loss.backward()
    return loss


def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n


for _ in range(3):
    loss = step()
_lib.reset_launch_count()
ms, wall = timed(step, 5)
print(f"train B={B}: {ms:.2f} ms/step (wall {wall:.2f}) -> {B / ms * 1e3:.1f} seg/s, loss {loss.item():.4f}, "
      f"launches/step {_lib.launch_count() / 5:.0f}, "
      f"{B * 164.502 * 3 / ms:.1f} TFLOP/s (3x fwd)")
m.eval()
hs = torch.randn(SB, 80, 862, device=dev)
zs = torch.randn(SB, 862 * 256, device=dev) * 0.6


def synth():
    with torch.no_grad():
        return m.infer(hs, 0.6, z=zs)


for _ in range(2):
    a = synth()
_lib.reset_launch_count()
ms, wall = timed(synth, 3)
print(f"synth B={SB}: {ms:.2f} ms (wall {wall:.2f}) -> {SB * 862 * 256 / ms:.0f} kHz, launches {_lib.launch_count() / 3:.0f}, "
      f"finite {torch.isfinite(a).all().item()}, {SB * 862 * 256 * 10.2814e6 / ms / 1e9:.1f} TFLOP/s")
