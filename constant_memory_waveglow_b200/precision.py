"""Operand precision of the WN GEMMs.

  fp32  exact CUDA-core engine (FFMA).  Selected automatically when TF32 is disabled for convolutions
        (``torch.backends.cudnn.allow_tf32 = False``) -- that is what the reference's tests
        (``tests/test_fwd_bwd.py:10-11``) and ``train.py --no-tf32`` (``train.py:92-97``) do to ask for
        full precision -- and for WN shapes the tensor-core engine does not tile.
  fp16  tcgen05 tensor cores (kind::f16), fp16 operands, fp32 accumulation in TMEM.  The default otherwise:
        fp16 has the SAME 10 mantissa bits as TF32 -- the reference's own default for convolutions on
        Ampere and later -- at twice TF32's tensor-core rate and half its operand bytes.  What fp16 lacks
        is exponent range; the forward activations of a weight-normed WN are O(1..100), and the backward
        chain, which is linear in the incoming cotangent, runs on S * cotangent with a power-of-two S picked
        per call on the device (``csrc/wn_kernels.cuh::grad_scale_kernel``) and is unscaled exactly where
        results leave the 16-bit slabs.  Measured against the fp32 oracle at the LJ config
        (profiles/r02_precision.json): z, logdet, audio and every gradient within 1e-3 rel-L2.
  bf16  same engine with bf16 operands (7 mantissa bits; rel-L2 1.4e-3 on z and up to 5e-3 on single
        gradient tensors at the LJ config, profiles/r01_precision.json): opt-in only.

Flow state, 1x1 convolutions, coupling arithmetic, `end` conv and log-determinants are always fp32.
Override with ``set_precision('fp32'|'bf16'|'fp16'|'auto')`` or the CMWG_PRECISION environment variable.
"""
from __future__ import annotations

import os

import torch

_VALID = ("auto", "fp32", "bf16", "fp16")
_mode = os.environ.get("CMWG_PRECISION", "auto").lower()
if _mode not in _VALID:
    raise ValueError(f"CMWG_PRECISION={_mode!r} not in {_VALID}")


def set_precision(mode: str) -> None:
    global _mode
    mode = mode.lower()
    if mode not in _VALID:
        raise ValueError(f"precision {mode!r} not in {_VALID}")
    _mode = mode


def get_precision() -> str:
    return _mode


def resolve(tc_supported: bool, training: bool) -> str:
    """Concrete precision for one WN call."""
    mode = _mode
    if mode == "auto":
        if not torch.backends.cudnn.allow_tf32:
            mode = "fp32"
        else:
            mode = "fp16"
    if mode != "fp32" and not tc_supported:
        mode = "fp32"
    return mode
