import torch, sys
a = torch.load("gpurun_out/mega_probe_mega.pt"); b = torch.load("gpurun_out/mega_probe_base.pt")
for k in a:
    print(k, "bit-identical" if torch.equal(a[k], b[k]) else f"DIFF max {(a[k].float() - b[k].float()).abs().max().item():.3e}")
