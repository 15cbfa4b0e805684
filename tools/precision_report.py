"""Parity report of the CUDA path against the CPU oracle at the LJ configuration (B=2, T=16000):
rel-L2 / max-abs of z, log-det, synthesis audio, parameter gradients (aggregate + worst tensor) and
the invertibility round trip, per operand precision.  Writes gpurun_out/precision.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import constant_memory_waveglow_b200 as cm  # noqa: E402
from constant_memory_waveglow_b200 import precision  # noqa: E402
from oracle import flow_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def main():
    torch.set_num_threads(os.cpu_count())
    spec = O.WaveGlowSpec(12, 8, 4, 2, 256, 80)
    B, T = int(os.environ.get("CMWG_REPORT_B", "2")), 16000
    out = {}
    for init, end_std in (("default_conv_init", None), ("end_std_0.05", 0.05)):
        sd = O.random_state(spec, 256, 8, seed=0, end_std=end_std)
        g = torch.Generator().manual_seed(0)
        x = torch.rand(B, T, generator=g) * 2 - 1
        h = torch.randn(B, 80, 63, generator=g)
        zs = torch.randn(B, 63 * 256, generator=g) * 0.6
        z_ref, ld_ref, loss_ref, grads_ref = O.waveglow_train_step(sd, spec, x, h, 0.7)
        audio_ref = O.waveglow_infer(sd, spec, h, zs)
        xr_ref, _ = O.waveglow_reverse(sd, spec, z_ref, h)
        res = {"oracle_roundtrip_max": (xr_ref - x).abs().max().item(), "oracle_roundtrip_rel_l2": rel(xr_ref, x)}
        for prec in ("fp32", "bf16", "fp16"):
            precision.set_precision(prec)
            m = cm.WaveGlow(12, 8, 4, 2, 256, 80, True, zero_init=False)
            m.load_state_dict(sd)
            m = m.cuda().train()
            r = {}
            if True:
                z, ld = m(x.cuda(), h.cuda())
                loss = cm.WaveGlowLoss(0.7)(z, ld)
                loss.backward()
                r["z_rel_l2"] = rel(z, z_ref)
                r["z_max_abs"] = (z.cpu() - z_ref).abs().max().item()
                r["logdet_rel"] = rel(ld, ld_ref)
                r["loss_rel"] = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
                num = den = 0.0
                worst = ("", 0.0)
                for n, p in m.named_parameters():
                    gr = grads_ref[n].double()
                    e2 = (p.grad.double().cpu() - gr).pow(2).sum().item()
                    num += e2
                    den += gr.pow(2).sum().item()
                    e = (e2 / max(gr.pow(2).sum().item(), 1e-300)) ** 0.5
                    if e > worst[1]:
                        worst = (n, e)
                r["grad_rel_l2_aggregate"] = (num / den) ** 0.5
                r["grad_rel_l2_worst"] = {"tensor": worst[0], "rel_l2": worst[1]}
            with torch.no_grad():
                m.eval()
                z2, ld2 = m(x.cuda(), h.cuda())
                r["z_rel_l2_nograd"] = rel(z2, z_ref)
                xr, _ = m.reverse(z2.clone(), h.cuda())
                r["roundtrip_max_abs"] = (xr.cpu() - x).abs().max().item()
                r["roundtrip_rel_l2"] = rel(xr, x)
                audio = m.infer(h.cuda(), 0.6, z=zs.cuda())
                r["audio_rel_l2"] = rel(audio, audio_ref)
            res[prec] = r
            del m
        out[init] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
