import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import constant_memory_waveglow_b200 as cm
from constant_memory_waveglow_b200 import precision
from oracle import flow_oracle as O
from tests._util import prefixed, to_double
from tests.test_gpu_wn import make_block

def run(cin, aux, ch, depth, B, T, direction, bias=False, reps=6):
    precision.set_precision("bf16")
    blk = make_block(cin, aux, ch, depth, True, seed=cin + depth, bias=bias)
    sd = prefixed(blk.state_dict(), "")
    sd64 = to_double({k: v.clone() for k, v in sd.items()})
    g = torch.Generator().manual_seed(100 + T)
    x = torch.rand(B, 2 * cin, T, generator=g, dtype=torch.float64) * 2 - 1
    y = torch.randn(B, aux, T, generator=g, dtype=torch.float64)
    dz = torch.randn(B, 2 * cin, T, generator=g, dtype=torch.float64) / (B * T)
    dls = torch.full((B, cin, T), -1.0 / (B * T), dtype=torch.float64)
    rev = direction == "reverse"
    out_ref, ls_ref, dx_ref, dp_ref, dy_ref = O.coupling_grads(sd64, "F.", x, y, dz, dls, reverse=rev, need_dy=True)
    blk = blk.cuda()
    res = []
    for rep in range(reps):
        xg = x.float().cuda().requires_grad_(True)
        yg = y.float().cuda().requires_grad_(True)
        xin = xg.clone()
        out, ls = (blk.reverse(xin, yg) if rev else blk(xin, yg))
        obj = (out * dz.float().cuda()).sum() + (ls * dls.float().cuda()).sum()
        blk.zero_grad()
        obj.backward()
        torch.cuda.synchronize()
        dy = yg.grad.double().cpu()
        res.append(((dy - dy_ref).norm() / dy_ref.norm()).item())
    print(f"[{direction} ch={ch} T={T}] dy rel:", " ".join(f"{r:.4f}" for r in res), flush=True)

print("env:", {k: v for k, v in os.environ.items() if k.startswith("CMWG")})
run(4, 12, 64, 2, 2, 300, "forward")
run(8, 40, 128, 4, 2, 4000, "reverse", reps=4)
