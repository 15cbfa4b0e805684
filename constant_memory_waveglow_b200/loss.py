"""Negative log-likelihood of the flow (reference ``model/loss.py:10-15``) as one fused kernel pair:
loss = mean_b(0.5 * sum_t z^2 / sigma^2 - logdet_b) [/ T]; the backward seeds the reversible
backward with dz = z / (sigma^2 B T) and dlogdet = -1 / (B T)."""
import torch

from . import ops


class _NLLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, logdet, sigma, mean):
        need = z.requires_grad or logdet.requires_grad
        loss, dz = ops.nll_loss(z.detach(), logdet.detach(), sigma, mean, want_dz=need)
        ctx.save_for_backward(dz)
        ctx.B, ctx.T, ctx.mean = z.shape[0], z.shape[1], mean
        ctx.logdet_shape = logdet.shape
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (dz,) = ctx.saved_tensors
        gz = dz * gloss if ctx.needs_input_grad[0] else None
        gl = None
        if ctx.needs_input_grad[1]:
            k = -1.0 / ctx.B / (ctx.T if ctx.mean else 1)
            gl = (gloss * k).expand(ctx.B).reshape(-1)
            if len(ctx.logdet_shape) == 0:
                gl = gl.sum()
            else:
                gl = gl.contiguous()
        return gz, gl, None, None


class WaveGlowLoss(torch.nn.Module):
    def __init__(self, sigma=1., elementwise_mean=True):
        super().__init__()
        self.sigma2 = sigma ** 2
        self.sigma = float(sigma)
        self.mean = elementwise_mean

    def forward(self, z, logdet):
        if z.dim() != 2:
            z = z.reshape(z.shape[0], -1)
        return _NLLFunction.apply(z, logdet, self.sigma, self.mean)
