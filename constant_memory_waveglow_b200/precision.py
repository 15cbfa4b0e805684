"""Operand precision of the WN GEMMs.

  fp32  exact CUDA-core engine (FFMA).  Selected automatically when TF32 is disabled for convolutions
        (``torch.backends.cudnn.allow_tf32 = False``) -- that is what the reference's tests
        (``tests/test_fwd_bwd.py:10-11``) and ``train.py --no-tf32`` (``train.py:92-97``) do to ask for
        full precision -- and for WN shapes the tensor-core engine does not tile.
  bf16  tcgen05 tensor cores, bf16 operands, fp32 accumulation in TMEM (default otherwise; the
        reference's own default on Ampere+ GPUs is TF32 convolutions).
  fp16  tcgen05 tensor cores, fp16 operands: 3 more mantissa bits than bf16 at the same speed;
        forward / synthesis only (gradients need bf16's exponent range).  ``auto`` picks it for calls
        that build no autograd graph (synthesis, evaluation): measured audio rel-L2 vs the fp32 oracle
        1.8e-4 against 1.4e-3 with bf16 operands (profiles/r01_precision.json).

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
            mode = "bf16" if training else "fp16"
    if mode != "fp32" and not tc_supported:
        mode = "fp32"
    if mode == "fp16" and training:
        mode = "bf16"
    return mode
