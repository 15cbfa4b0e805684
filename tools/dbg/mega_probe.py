"""Single-kernel WN forward (CMWG_MEGA) vs the layer-at-a-time pipeline: writes outputs for a cross-process bit
comparison, prints rel-L2 against the fp32 engine and timings."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm
tag = sys.argv[1]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
wn = cm.WN(4, 80, zero_init=False).to(dev)
out = {}
for (B, T) in ((2, 700), (3, 2000), (24, 2000)):
    x = torch.randn(B, 8, T, device=dev)
    y = torch.randn(B, 80, T, device=dev)
    ref, _ = wn._cmwg_forward(x, y, save=False, prec="fp32")
    for prec, save in (("bf16", False), ("bf16", True), ("fp16", False)):
        lst, st = wn._cmwg_forward(x, y, save=save, prec=prec)
        torch.cuda.synchronize()
        rel = ((lst - ref).norm() / ref.norm()).item()
        out[f"{B}x{T}_{prec}_{int(save)}"] = lst.cpu()
        if save:
            out[f"{B}x{T}_saved"] = st.saved.cpu()[:1 << 22].clone()
        print(f"{tag} B={B} T={T} {prec} save={save}: rel-L2 vs fp32 engine {rel:.3e} finite={bool(torch.isfinite(lst).all())}", flush=True)
torch.save(out, os.path.join(ROOT, "gpurun_out", f"mega_probe_{tag}.pt"))
x = torch.randn(24, 8, 2000, device=dev); y = torch.randn(24, 80, 2000, device=dev)
for save in (False, True):
    for _ in range(3): wn._cmwg_forward(x, y, save=save, prec="bf16")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): wn._cmwg_forward(x, y, save=save, prec="bf16")
    b.record(); torch.cuda.synchronize()
    print(f"{tag} forward B=24 save={save}: {a.elapsed_time(b) / 20:.3f} ms", flush=True)
