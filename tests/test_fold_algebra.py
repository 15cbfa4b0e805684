"""The identities behind the folded start / end convs of the task kernels (csrc/wn_kernels.cuh: pack_fold0_kernel,
fold0_dw_kernel, pack_foldend_kernel, foldend_dw_kernel), checked in fp64 on the CPU against the reference's own formulation
(model/waveglow.py:74,92,98-105).  The CUDA path is held to the oracle by the GPU tests; this file pins the algebra the
kernels implement, independent of any device."""
import torch
import torch.nn.functional as F


def _setup(cin=4, C=16, T=37, B=2, R=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)   # noqa: E731
    return dict(xa=r(B, cin, T), Wstart=r(C, cin, 1), W0=r(2 * C, C, R), Wo=r(2 * C, C, 1), Wend=r(2 * cin, C, 1),
                g0=r(B, C, T), dpre0=r(B, 2 * C, T), dlst=r(B, 2 * cin, T), cin=cin, C=C, T=T, B=B, R=R)


def test_start_conv_folds_into_layer_0():
    """h_0 = W_start x_a (no bias), layer 0's dilated conv (dilation 1, 'same' zero padding of h_0):
    conv(W_0, h_0)[t] = sum_tap (W_0,tap W_start) x_a[t + tap - 1] with x_a zero outside [0, T) -- what PA0f holds, against the
    taps cond_aug_kernel writes; and h_1 = W_res g_0 + W_start x_a (PB0f)."""
    s = _setup()
    h0 = F.conv1d(s["xa"], s["Wstart"])
    ref = F.conv1d(h0, s["W0"], padding=1)
    Wf = torch.einsum("okt,kc->oct", s["W0"], s["Wstart"][:, :, 0])            # [2C][cin][R]
    # taps[b, c, j, t] = x_a[b, c, t + j - 1], zero outside [0, T): the columns cond_aug_kernel writes
    taps = torch.zeros(s["B"], s["cin"], 3, s["T"], dtype=torch.float64)
    for j in range(3):
        sh = j - 1
        lo, hi = max(0, -sh), min(s["T"], s["T"] - sh)
        taps[:, :, j, lo:hi] = s["xa"][:, :, lo + sh:hi + sh]
    folded = torch.einsum("oct,bctn->bon", Wf, taps)
    assert torch.allclose(folded, ref, rtol=1e-12, atol=1e-12)
    res = F.conv1d(s["g0"], s["Wo"][:s["C"]])
    assert torch.allclose(res + torch.einsum("kc,bct->bkt", s["Wstart"][:, :, 0], s["xa"]), res + h0, rtol=1e-12, atol=1e-12)


def test_layer_0_weight_gradient_through_the_fold():
    """dW_0[o][k][tap] = sum_t dpre_0[t][o] h_0[t + tap - 1][k] = sum_c D[o][tap][c] W_start[k][c] with
    D[o][tap][c] = sum_t dpre_0[t][o] x_a[t + tap - 1][c] -- the columns of the conditioning weight-gradient tile that face
    the x_a taps (fold0_dw_kernel)."""
    s = _setup(seed=1)
    W0 = s["W0"].clone().requires_grad_(True)
    h0 = F.conv1d(s["xa"], s["Wstart"])
    (F.conv1d(h0, W0, padding=1) * s["dpre0"]).sum().backward()
    xp = F.pad(s["xa"], (1, 1))
    D = torch.stack([torch.einsum("bot,bct->oc", s["dpre0"], xp[..., j:j + s["T"]]) for j in range(3)], 1)   # [2C][R][cin]
    dW0 = torch.einsum("ojc,kc->okj", D, s["Wstart"][:, :, 0])
    assert torch.allclose(dW0, W0.grad, rtol=1e-11, atol=1e-11)


def test_end_conv_folds_into_the_backward():
    """lst = W_end sum_i W_skip,i g_i.  With dskip = W_end^T dlst:  W_skip^T dskip = (W_end W_skip)^T dlst (the dgate tiles'
    folded k-block, Q1f);  dW_skip = dskip^T-outer-g = W_end^T P,  dW_end = P W_skip^T-sum  with  P[o][n] = sum_t dlst[t][o] g[t][n]
    (foldend_dw_kernel)."""
    s = _setup(seed=2)
    C = s["C"]
    Wskip = s["Wo"][C:].clone().requires_grad_(True)                    # [Cs = C][Cd = C][1]
    Wend = s["Wend"].clone().requires_grad_(True)
    g0 = s["g0"].clone().requires_grad_(True)
    lst = F.conv1d(F.conv1d(g0, Wskip), Wend)
    (lst * s["dlst"]).sum().backward()
    We, Ws = s["Wend"][:, :, 0], s["Wo"][C:, :, 0]                       # [cout][Cs], [Cs][Cd]
    dskip = torch.einsum("ok,bot->bkt", We, s["dlst"])
    dg_ref = torch.einsum("kn,bkt->bnt", Ws, dskip)
    Ffold = (We @ Ws)                                                    # [cout][Cd] = (W_end W_skip)
    assert torch.allclose(torch.einsum("on,bot->bnt", Ffold, s["dlst"]), dg_ref, rtol=1e-11, atol=1e-11)
    assert torch.allclose(dg_ref, g0.grad, rtol=1e-11, atol=1e-11)
    P = torch.einsum("bot,bnt->on", s["dlst"], s["g0"])                  # [cout][Cd]
    assert torch.allclose(We.t() @ P, Wskip.grad[:, :, 0], rtol=1e-11, atol=1e-11)
    assert torch.allclose(P @ Ws.t(), Wend.grad[:, :, 0], rtol=1e-11, atol=1e-11)


def test_tanh_recovered_from_gate_and_sigmoid():
    """The saving forward stores g = tanh * sigmoid and the sigmoid; the gate backward uses tanh = g / sigmoid
    (GateBwdTcEpi): both gradient halves of fused_gate (model/waveglow.py:13-15) in those terms."""
    g = torch.Generator().manual_seed(3)
    a = torch.randn(1000, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(1000, generator=g, dtype=torch.float64, requires_grad=True)
    dg = torch.randn(1000, generator=g, dtype=torch.float64)
    (torch.tanh(a) * torch.sigmoid(b) * dg).sum().backward()
    s_ = torch.sigmoid(b.detach())
    gate = torch.tanh(a.detach()) * s_
    t = gate / s_
    assert torch.allclose(dg * s_ * (1 - t * t), a.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dg * t * s_ * (1 - s_), b.grad, rtol=1e-10, atol=1e-12)
