"""Shared helpers of the test-suite (CPU oracle side + CUDA product side)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def to_double(sd):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}


def wn_kwargs(wn_ch, depth, **extra):
    kw = dict(dilation_channels=wn_ch, residual_channels=wn_ch, skip_channels=wn_ch, depth=depth)
    kw.update(extra)
    return kw


def prefixed(sd, prefix):
    return {prefix + k: v for k, v in sd.items()}


# tolerances (rel-L2 against the fp64 oracle) per operand precision of the WN GEMMs
TOL = {
    "fp32": dict(out=2e-6, logdet=1e-5, grad=2e-5, roundtrip=5e-6),
    "fp16": dict(out=2e-3, logdet=2e-3, grad=None, roundtrip=5e-3),
    "bf16": dict(out=1e-2, logdet=1e-2, grad=3e-2, roundtrip=2e-2),
}
