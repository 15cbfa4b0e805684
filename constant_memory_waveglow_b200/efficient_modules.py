"""Reversible flow primitives with constant-memory backward: host-side mirror of the reference's
``model/efficient_modules.py`` (``InvertibleConv1x1`` :17-54, ``AffineCouplingBlock`` :57-96,
``AffineCouplingFunc`` :99-154, ``InvAffineCouplingFunc`` :157-212, ``Conv1x1Func`` :215-244,
``InvConv1x1Func`` :247-279).

Contract kept from the reference:
  * memory-efficient mode CONSUMES the input: after the op the module frees the input tensor's
    storage (``resize_(0)``); the backward re-materialises the input from the saved OUTPUT into the
    very same storage object, which revives the upstream op's saved output (its output *is* that
    tensor) -- so activations are never stored, only recomputed flow by flow;
  * ``Func.apply(x, y, F, *F.parameters())`` / ``Func.apply(x, weight)`` signatures, gradient order
    equal to ``F.parameters()`` order, ``dy`` only when the conditioning needs a gradient;
  * forward returns ``(z, log_s)`` / ``(z, logdet)``, reverse returns ``(x, -log_s)`` / ``(x, -logdet)``,
    bitwise deterministic.

All arithmetic runs in libcmwg_b200.so; when the transform ``F`` is the package's WN the whole
recompute + gradient pipeline is fused (``F._cmwg_forward`` / ``F._cmwg_backward``); any other
``transform_type`` is called as a module and differentiated with autograd (generic path).
"""
from __future__ import annotations

import threading
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor
from torch.autograd import Function

from . import _lib as L
from . import ops
from .base import Reversible

__all__ = ['InvertibleConv1x1', 'AffineCouplingBlock']

try:  # torch >= 2.4
    from torch.amp import custom_bwd as _cb, custom_fwd as _cf

    def custom_fwd(fn):
        return _cf(fn, device_type='cuda')

    def custom_bwd(fn):
        return _cb(fn, device_type='cuda')
except ImportError:  # pragma: no cover
    from torch.cuda.amp import custom_bwd, custom_fwd


# ---------------------------------------------------------------------------------------------
# Inside Function.forward grad mode is always off and ctx.needs_input_grad only mirrors
# requires_grad flags, so the module wrappers record whether a graph is being built at all
# (torch.no_grad() synthesis must neither keep activations nor be forced onto the training precision).
# ---------------------------------------------------------------------------------------------
_tls = threading.local()


class grad_hint:
    def __enter__(self):
        self.prev = getattr(_tls, "grad", None)
        _tls.grad = torch.is_grad_enabled()

    def __exit__(self, *exc):
        _tls.grad = self.prev


def graph_possible() -> bool:
    g = getattr(_tls, "grad", None)
    return True if g is None else g


# ---------------------------------------------------------------------------------------------
# storage free / restore (the constant-memory mechanism)
# ---------------------------------------------------------------------------------------------
def _free_storage(t: Tensor) -> None:
    t.untyped_storage().resize_(0)


def _restore_storage(t: Tensor) -> None:
    need = (t.storage_offset() + (t.numel() if t.is_contiguous() else _span(t))) * t.element_size()
    st = t.untyped_storage()
    if st.size() < need:
        st.resize_(need)


def _span(t: Tensor) -> int:
    return 1 + sum((s - 1) * st for s, st in zip(t.shape, t.stride()) if s > 0)


def _write_restored(dst: Tensor, src_fn) -> None:
    """Re-materialise into `dst`'s own storage; `src_fn(out)` fills a contiguous (B, C, T) tensor."""
    _restore_storage(dst)
    if dst.is_contiguous():
        src_fn(dst)
    else:
        tmp = torch.empty(dst.shape, device=dst.device, dtype=dst.dtype)
        src_fn(tmp)
        dst.data.copy_(tmp)


# ---------------------------------------------------------------------------------------------
# invertible 1x1 convolution
# ---------------------------------------------------------------------------------------------
def _conv1x1_fwd(ctx, x, weight, inverse: bool):
    L.require_cuda(x, weight, op="InvertibleConv1x1")
    w2 = weight.detach().reshape(weight.shape[0], weight.shape[1]).float().contiguous()
    T = x.shape[-1]
    winv, logdet = ops.small_inverse_logdet(w2)
    xd = x.detach()
    z = ops.conv1x1_apply(winv if inverse else w2, xd)
    log_det_w = (-logdet if inverse else logdet) * T
    ctx.inverse = inverse
    ctx.save_for_backward(xd, weight, z, winv)
    return z, log_det_w


def _conv1x1_bwd(ctx, z_grad, log_det_grad, restore: bool):
    x, weight, z, winv = ctx.saved_tensors          # winv: W^-1 from the forward pass (the weight has not changed since)
    inverse = ctx.inverse
    w2 = weight.detach().reshape(weight.shape[0], weight.shape[1]).float().contiguous()
    z_grad = z_grad.contiguous() if z_grad is not None else torch.zeros_like(z)
    dw = None
    if ctx.needs_input_grad[1]:
        # written straight into the parameter's slot of its data-parallel gradient bucket when there is one (parallel.py):
        # autograd then adopts the view as .grad without a copy kernel
        from .parallel import grad_buffer
        dw = grad_buffer(weight)
    # forward was z = M x with M = W (or W^-1): x = M^-1 z goes back into the freed storage (restore) or into a scratch
    # tensor (stored mode: only the weight gradient needs it); dx = M^T dz and dW come from the same sweep
    direct = restore and x.is_contiguous()
    if restore:
        _restore_storage(x)
    dst = x if direct else (torch.empty(z.shape, device=z.device) if (restore or dw is not None) else None)
    dx = ops.conv1x1_backward(w2, winv, inverse, z, z_grad, log_det_grad, dst,
                              None if dw is None else dw.view(w2.shape))
    if restore and not direct:
        x.data.copy_(dst)
    return dx, dw


class Conv1x1Func(Function):
    """z = W x, logdet = T log det W; backward restores x = W^-1 z into the freed input storage
    (reference ``model/efficient_modules.py:215-244``)."""

    @staticmethod
    @custom_fwd
    def forward(ctx, x, weight):
        return _conv1x1_fwd(ctx, x, weight, inverse=False)

    @staticmethod
    @custom_bwd
    def backward(ctx, z_grad, log_det_W_grad):
        return _conv1x1_bwd(ctx, z_grad, log_det_W_grad, restore=True)


class InvConv1x1Func(Function):
    """z = W^-1 x, logdet = -T log det W (reference ``model/efficient_modules.py:247-279``)."""

    @staticmethod
    @custom_fwd
    def forward(ctx, x, inv_weight):
        return _conv1x1_fwd(ctx, x, inv_weight, inverse=True)

    @staticmethod
    @custom_bwd
    def backward(ctx, z_grad, log_det_W_grad):
        return _conv1x1_bwd(ctx, z_grad, log_det_W_grad, restore=True)


class _StoredConv1x1Func(Function):
    """Same kernels, input kept alive (what the reference's naive mode gets from plain autograd)."""

    @staticmethod
    @custom_fwd
    def forward(ctx, x, weight, inverse):
        return _conv1x1_fwd(ctx, x, weight, inverse=inverse)

    @staticmethod
    @custom_bwd
    def backward(ctx, z_grad, log_det_W_grad):
        return _conv1x1_bwd(ctx, z_grad, log_det_W_grad, restore=False) + (None,)


class InvertibleConv1x1(Reversible, nn.Conv1d):
    """Reference ``model/efficient_modules.py:17-54``: an ``nn.Conv1d(c, c, 1, bias=False)`` whose
    weight is initialised to a random rotation with positive determinant."""

    def __init__(self, c, memory_efficient=False, reverse_mode=False):
        super().__init__(in_channels=c, out_channels=c, kernel_size=1, bias=False, reverse_mode=reverse_mode)
        q = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        self.weight.data[:] = q.contiguous().unsqueeze(-1)
        if memory_efficient:
            self._efficient_forward = Conv1x1Func.apply
            self._efficient_reverse = InvConv1x1Func.apply

    def forward_computation(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        with grad_hint():
            if hasattr(self, '_efficient_forward'):
                z, log_det_w = self._efficient_forward(x, self.weight)
                _free_storage(x)
                return z, log_det_w
            return _StoredConv1x1Func.apply(x, self.weight, False)

    def reverse_computation(self, z: Tensor) -> Tuple[Tensor, Tensor]:
        with grad_hint():
            if hasattr(self, '_efficient_reverse'):
                x, log_det_w = self._efficient_reverse(z, self.weight)
                _free_storage(z)
                return x, log_det_w
            return _StoredConv1x1Func.apply(z, self.weight, True)


# ---------------------------------------------------------------------------------------------
# affine coupling
# ---------------------------------------------------------------------------------------------
def _is_fused(F) -> bool:
    return bool(getattr(F, "_cmwg_fused", False))


def _check_coupling_input(x: Tensor) -> None:
    if x.dim() != 3 or x.shape[1] % 2:
        raise ValueError(f"affine coupling expects (B, even C, T), got {tuple(x.shape)}")


def _coupling_fwd(ctx, x, y, F, inverse: bool, recompute: bool):
    """Shared forward of the four coupling Functions.

    recompute=True : constant-memory mode, nothing but (x, y, out) is kept; WN runs without saving.
    recompute=False: activation-storing mode (naive), WN saves its per-layer activations.
    """
    L.require_cuda(x, y, op="AffineCoupling")
    _check_coupling_input(x)
    xd = ops._ncl(x.detach())
    yd = y.detach()
    ctx.F, ctx.inverse, ctx.recompute = F, inverse, recompute
    need = any(ctx.needs_input_grad) and graph_possible()
    if _is_fused(F):
        from . import precision
        prec = precision.resolve(F._tc_supported(), training=need)
        lst, st = F._cmwg_forward(xd, yd, save=(need and not recompute), prec=prec)
        ctx.prec = prec
        ctx.st = None if recompute else st
    else:
        cin = xd.shape[1] // 2
        with torch.no_grad():
            log_s, t = F(xd[:, :cin].contiguous(), yd)
        lst = torch.cat((log_s, t), 1).float().contiguous()
        ctx.st = None
        ctx.prec = None
    out, neg = ops.coupling_apply(xd, lst, inverse)
    cin = xd.shape[1] // 2
    # keep an ALIAS of the caller's tensor (same storage object): the backward re-materialises the
    # input into exactly that storage, which is what revives the upstream op's saved output
    x_alias = x.detach()
    if recompute:
        ctx.save_for_backward(x_alias, y, out)
    else:
        ctx.save_for_backward(x_alias, y, out, lst)
    return out, (neg if inverse else lst[:, :cin])


def _coupling_bwd(ctx, out_grad, ls_grad):
    F, inverse, recompute = ctx.F, ctx.inverse, ctx.recompute
    saved = ctx.saved_tensors
    x, y, out = saved[0], saved[1], saved[2]
    B, c, T = out.shape
    cin = c // 2
    need_dy = ctx.needs_input_grad[1]
    if out_grad is None:
        out_grad = torch.zeros_like(out)
    if ls_grad is None:
        ls_grad = torch.zeros((B, cin, T), device=out.device)
    fused = _is_fused(F)

    # 1. (re)compute the WN output from the untouched half of the saved OUTPUT (xa == za)
    if recompute:
        if fused:
            lst, st = F._cmwg_forward(out, y.detach(), save=True, prec=ctx.prec, transient=True)
        else:
            xa = out[:, :cin].detach().contiguous().requires_grad_(True)
            yy = y.detach().requires_grad_(need_dy)
            with torch.enable_grad():
                log_s, t = F(xa, yy)
                lst_graph = torch.cat((log_s, t), 1)
            lst = lst_graph.detach().float().contiguous()
    else:
        lst = saved[3]
        st = ctx.st
        if not fused:
            xa = x[:, :cin].detach().contiguous().requires_grad_(True)
            yy = y.detach().requires_grad_(need_dy)
            with torch.enable_grad():
                log_s, t = F(xa, yy)
                lst_graph = torch.cat((log_s, t), 1)

    # 2. restore the input into its own (freed) storage + elementwise cotangents
    if recompute:
        _restore_storage(x)
        restored = x if x.is_contiguous() else torch.empty(out.shape, device=out.device)
    else:
        restored = torch.empty(out.shape, device=out.device)
    dlst, din = ops.coupling_bwd(out, lst, out_grad, ls_grad, restored, inverse)
    if recompute and restored is not x:
        x.data.copy_(restored)

    # 3. gradient through the transform
    if fused:
        grads, dy = F._cmwg_backward(st, out, dlst, din, need_dy)
        ctx.st = None
    else:
        wrt = [xa] + list(F.parameters()) + ([yy] if need_dy else [])
        g = torch.autograd.grad(lst_graph, wrt, grad_outputs=dlst, allow_unused=True)
        din[:, :cin] += g[0]
        n = len(wrt) - 1 - (1 if need_dy else 0)
        grads = list(g[1:1 + n])
        dy = g[-1] if need_dy else None
    return (din, dy, None) + tuple(grads)


class AffineCouplingFunc(Function):
    """za = xa, zb = xb*exp(log_s) + t with (log_s, t) = F(xa, y); constant-memory backward
    (reference ``model/efficient_modules.py:99-154``)."""

    @staticmethod
    @custom_fwd
    def forward(ctx, x, y, F, *F_weights):
        return _coupling_fwd(ctx, x, y, F, inverse=False, recompute=True)

    @staticmethod
    @custom_bwd
    def backward(ctx, z_grad, log_s_grad):
        return _coupling_bwd(ctx, z_grad, log_s_grad)


class InvAffineCouplingFunc(Function):
    """xa = za, xb = (zb - t)/exp(log_s), returns (x, -log_s); constant-memory backward
    (reference ``model/efficient_modules.py:157-212``)."""

    @staticmethod
    @custom_fwd
    def forward(ctx, z, y, F, *F_weights):
        return _coupling_fwd(ctx, z, y, F, inverse=True, recompute=True)

    @staticmethod
    @custom_bwd
    def backward(ctx, x_grad, log_s_grad):
        return _coupling_bwd(ctx, x_grad, log_s_grad)


class _StoredCouplingFunc(Function):
    """Activation-storing variant (same kernels) used when ``memory_efficient=False``."""

    @staticmethod
    @custom_fwd
    def forward(ctx, x, y, F, inverse, *F_weights):
        return _coupling_fwd(ctx, x, y, F, inverse=inverse, recompute=False)

    @staticmethod
    @custom_bwd
    def backward(ctx, out_grad, ls_grad):
        r = _coupling_bwd(ctx, out_grad, ls_grad)
        return r[:3] + (None,) + r[3:]


class AffineCouplingBlock(Reversible):
    """Reference ``model/efficient_modules.py:57-96``: owns ``self.F = transform_type(**kwargs)``."""

    def __init__(self, transform_type, memory_efficient=True, reverse_mode=False, **kwargs):
        super().__init__(reverse_mode)
        self.F = transform_type(**kwargs)
        if memory_efficient:
            self._efficient_forward = AffineCouplingFunc.apply
            self._efficient_reverse = InvAffineCouplingFunc.apply

    def forward_computation(self, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
        with grad_hint():
            if hasattr(self, '_efficient_forward'):
                z, log_s = self._efficient_forward(x, y, self.F, *self.F.parameters())
                _free_storage(x)
                return z, log_s
            return _StoredCouplingFunc.apply(x, y, self.F, False, *self.F.parameters())

    def reverse_computation(self, z: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
        with grad_hint():
            if hasattr(self, '_efficient_reverse'):
                x, log_s = self._efficient_reverse(z, y, self.F, *self.F.parameters())
                _free_storage(z)
                return x, log_s
            return _StoredCouplingFunc.apply(z, y, self.F, True, *self.F.parameters())
