"""CPU oracle for the constant-memory-waveglow flow hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``constant_memory_waveglow_b200``) never imports anything from ``oracle/``.

What it is: a functional, state-dict driven restatement in plain fp32 (or fp64) PyTorch-on-CPU of
the reference algorithm for the path named in BASELINE.json ``north_star``.  Every function cites
the reference ``file:line`` (relative to the reference repo root) whose arithmetic it follows.
Weights are taken from a flat ``dict[str, Tensor]`` with the reference's state-dict keys
(``SURVEY.md`` section 8b), so the oracle, the reference and the CUDA product can share one set of
random-init weights.

Parity pin: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (CPU, fp32), runs it on seeded inputs in BOTH memory modes and stores
inputs / weights / outputs / gradients under ``tests/golden/*.pt``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.  The reference
itself ships no golden vectors (its tests are self-consistency checks), so the fixtures generated
from the reference are the pin.

Gradients: the reference's constant-memory backward (``model/efficient_modules.py:118-154,
176-212,230-244,263-279``) computes exactly the autograd gradient of the naive formulation; the
oracle therefore differentiates the naive restatement with ``torch.autograd`` and, separately,
restates the input reconstruction step (``coupling_restore_input`` / ``conv1x1_restore_input``)
that the reversible backward performs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# weight handling
# --------------------------------------------------------------------------------------------
def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v|| with the norm over every dim except 0 (``utils.py:14-16`` ->
    ``nn.utils.weight_norm`` default ``dim=0``)."""
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def resolve_weight(sd: State, prefix: str) -> Tensor:
    """Return the effective conv weight for ``prefix`` whether or not weight-norm is attached
    (``weight_g``/``weight_v`` pair vs plain ``weight`` after ``remove_weight_norms``,
    ``utils.py:9-16``)."""
    if prefix + "weight_g" in sd:
        return weight_norm_weight(sd[prefix + "weight_g"], sd[prefix + "weight_v"])
    return sd[prefix + "weight"]


def resolve_bias(sd: State, prefix: str) -> Optional[Tensor]:
    return sd.get(prefix + "bias")


def wn_depth(sd: State, prefix: str) -> int:
    n = 0
    while (prefix + f"layers.{n}.W.weight_g" in sd) or (prefix + f"layers.{n}.W.weight" in sd):
        n += 1
    return n


# --------------------------------------------------------------------------------------------
# WN transform (model/waveglow.py:13-105)
# --------------------------------------------------------------------------------------------
def fused_gate(x1: Tensor, x2: Tensor) -> Tensor:
    """``model/waveglow.py:13-15``."""
    return torch.tanh(x1) * torch.sigmoid(x2)


def wn_forward(sd: State, prefix: str, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """WN.forward (``model/waveglow.py:98-105``) with NonCausalLayer.forward (``:41-46``).

    x: (B, in_channels, T) ; y: (B, aux_channels, T) -> (log_s, t) each (B, in_channels, T).
    Dilation of layer i is 2**i and the padding is dilation*(radix-1)//2 (``:27,61``).
    """
    depth = wn_depth(sd, prefix)
    h = F.conv1d(x, resolve_weight(sd, prefix + "start."), resolve_bias(sd, prefix + "start."))
    v_all = F.conv1d(y, resolve_weight(sd, prefix + "V."), resolve_bias(sd, prefix + "V."))
    cum_skip = None
    for i, v in enumerate(v_all.chunk(depth, 1)):
        w = resolve_weight(sd, prefix + f"layers.{i}.W.")
        radix = w.shape[2]
        dil = 2 ** i
        xy = F.conv1d(h, w, resolve_bias(sd, prefix + f"layers.{i}.W."),
                      padding=dil * (radix - 1) // 2, dilation=dil) + v
        zw, zf = xy.chunk(2, 1)
        g = fused_gate(zw, zf)
        wo = resolve_weight(sd, prefix + f"layers.{i}.W_o.")
        ro = F.conv1d(g, wo, resolve_bias(sd, prefix + f"layers.{i}.W_o."))
        if i < depth - 1:
            res_ch = h.shape[1]
            h = ro[:, :res_ch] + h
            skip = ro[:, res_ch:]
        else:
            skip = ro
        cum_skip = skip if cum_skip is None else cum_skip + skip
    out = F.conv1d(cum_skip, sd[prefix + "end.weight"], sd.get(prefix + "end.bias"))
    log_s, t = out.chunk(2, 1)
    return log_s, t


# --------------------------------------------------------------------------------------------
# affine coupling (model/efficient_modules.py:57-212)
# --------------------------------------------------------------------------------------------
def coupling_forward(sd: State, prefix: str, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """``model/efficient_modules.py:77-82`` (naive) == ``:105-111`` (efficient forward)."""
    xa, xb = x.chunk(2, 1)
    log_s, t = wn_forward(sd, prefix, xa, y)
    zb = xb * log_s.exp() + t
    return torch.cat((xa, zb), 1), log_s


def coupling_reverse(sd: State, prefix: str, z: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """``model/efficient_modules.py:91-96`` == ``:163-172``; returns (x, -log_s)."""
    za, zb = z.chunk(2, 1)
    log_s, t = wn_forward(sd, prefix, za, y)
    xb = (zb - t) / log_s.exp()
    return torch.cat((za, xb), 1), -log_s


def coupling_restore_input(sd: State, prefix: str, z: Tensor, y: Tensor) -> Tensor:
    """Input reconstruction done inside the reversible backward
    (``model/efficient_modules.py:127-136``): x = cat(za, (zb - t)/exp(log_s))."""
    return coupling_reverse(sd, prefix, z, y)[0]


# --------------------------------------------------------------------------------------------
# invertible 1x1 convolution (model/efficient_modules.py:17-54, 215-279)
# --------------------------------------------------------------------------------------------
def conv1x1_forward(weight: Tensor, x: Tensor) -> Tuple[Tensor, Tensor]:
    """``model/efficient_modules.py:36-41`` / ``:218-226``: z = W x, logdet = T * logdet(W)."""
    n = x.shape[-1]
    return F.conv1d(x, weight), n * weight.squeeze(-1).logdet()


def conv1x1_reverse(weight: Tensor, z: Tensor) -> Tuple[Tensor, Tensor]:
    """``model/efficient_modules.py:49-54`` / ``:250-259``: x = W^-1 z, logdet = -T*logdet(W)."""
    n = z.shape[-1]
    w = weight.squeeze(-1)
    return F.conv1d(z, w.inverse().unsqueeze(-1)), -n * w.logdet()


def conv1x1_restore_input(weight: Tensor, z: Tensor) -> Tensor:
    """``model/efficient_modules.py:235-237``: x = W^-1 z written back into the freed storage."""
    return F.conv1d(z, weight.squeeze(-1).inverse().unsqueeze(-1))


def conv1x1_backward(weight: Tensor, x: Tensor, dz: Tensor, dlogdet: Tensor) -> Tuple[Tensor, Tensor]:
    """Closed form of ``Conv1x1Func.backward`` (``model/efficient_modules.py:230-244``):
    dx = W^T dz ; dW = sum_{b,t} dz x^T + W^-T * dlogdet * T."""
    n = x.shape[-1]
    w = weight.squeeze(-1)
    dx = F.conv1d(dz, w.t().unsqueeze(-1))
    dw = torch.einsum("bot,bit->oi", dz, x) + w.inverse().t() * dlogdet * n
    return dx, dw.unsqueeze(-1)


def invconv1x1_backward(weight: Tensor, x: Tensor, dz: Tensor, dlogdet: Tensor) -> Tuple[Tensor, Tensor]:
    """Closed form of ``InvConv1x1Func.backward`` (``model/efficient_modules.py:263-279``) where
    forward was z = W^-1 x, L = -T logdet W and ``x`` here is that forward's INPUT:
    dx = W^-T dz ; dM = sum dz (W z)^T = sum dz x^T ; dW = -W^-T dM W^-T - W^-T dlogdet T."""
    n = x.shape[-1]
    w = weight.squeeze(-1)
    wt_inv = w.inverse().t()
    dx = F.conv1d(dz, wt_inv.unsqueeze(-1))
    dm = torch.einsum("bot,bit->oi", dz, x)
    dw = -wt_inv @ dm @ wt_inv - wt_inv * dlogdet * n
    return dx, dw.unsqueeze(-1)


# --------------------------------------------------------------------------------------------
# WaveGlow model glue (model/waveglow.py:108-212), loss (model/loss.py:10-15), infer (base.py:42-55)
# --------------------------------------------------------------------------------------------
class WaveGlowSpec:
    """Structural hyper-parameters of ``WaveGlow.__init__`` (``model/waveglow.py:108-148``)."""

    def __init__(self, flows: int, n_group: int, n_early_every: int, n_early_size: int,
                 hop_size: int, n_mels: int):
        self.flows = flows
        self.n_group = n_group
        self.n_early_every = n_early_every
        self.n_early_size = n_early_size
        self.hop_size = hop_size
        self.n_mels = n_mels
        self.upsample_factor = hop_size // n_group
        self.sub_win = self.upsample_factor * 2 + 1
        self.up_pad = self.sub_win // 2 - self.upsample_factor // 2
        rem = n_group
        self.channels: List[int] = []
        self.z_split_sizes: List[int] = []
        for k in range(flows):
            if k % n_early_every == 0 and k:
                rem -= n_early_size
                self.z_split_sizes.append(n_early_size)
            self.channels.append(rem)
        self.z_split_sizes.append(rem)


def upsample_h(sd: State, spec: WaveGlowSpec, h: Tensor) -> Tensor:
    """``model/waveglow.py:126-130,210-212``: weight-normed depthwise ConvTranspose1d with bias."""
    w = resolve_weight(sd, "upsampler.")
    return F.conv_transpose1d(h, w, sd.get("upsampler.bias"), stride=spec.upsample_factor,
                              padding=spec.up_pad, groups=spec.n_mels)


def squeeze(x: Tensor, n_group: int) -> Tensor:
    """``model/waveglow.py:153``: (B, T) -> (B, n_group, T/n_group)."""
    return x.view(x.size(0), -1, n_group).transpose(1, 2).contiguous()


def unsqueeze(x: Tensor) -> Tensor:
    """``model/waveglow.py:179,207``: (B, C, T') -> (B, C*T')."""
    return x.transpose(1, 2).contiguous().view(x.size(0), -1)


def waveglow_forward(sd: State, spec: WaveGlowSpec, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``WaveGlow.forward_computation`` (``model/waveglow.py:150-179``)."""
    y = upsample_h(sd, spec, h)
    x = squeeze(x, spec.n_group)
    assert x.size(2) <= y.size(2)
    y = y[..., :x.size(2)]
    outs = []
    logdet = 0
    for k in range(spec.flows):
        if k % spec.n_early_every == 0 and k:
            outs.append(x[:, :spec.n_early_size])
            x = x[:, spec.n_early_size:]
        x, ldw = conv1x1_forward(sd[f"invconv1x1.{k}.weight"], x)
        x, log_s = coupling_forward(sd, f"WNs.{k}.F.", x, y)
        logdet = logdet + ldw + log_s.sum((1, 2))
    outs.append(x)
    return unsqueeze(torch.cat(outs, 1)), logdet


def waveglow_reverse(sd: State, spec: WaveGlowSpec, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``WaveGlow.reverse_computation`` (``model/waveglow.py:181-208``)."""
    y = upsample_h(sd, spec, h)
    z = squeeze(z, spec.n_group)
    assert z.size(2) <= y.size(2)
    y = y[..., :z.size(2)]
    *remained, z = z.split(spec.z_split_sizes, 1)
    remained = list(remained)
    logdet = 0
    for k in range(spec.flows - 1, -1, -1):
        z, log_s = coupling_reverse(sd, f"WNs.{k}.F.", z, y)
        z, ldw = conv1x1_reverse(sd[f"invconv1x1.{k}.weight"], z)
        logdet = logdet + ldw + log_s.sum((1, 2))
        if k % spec.n_early_every == 0 and k:
            z = torch.cat((remained.pop(), z), 1)
    return unsqueeze(z), logdet


def waveglow_loss(z: Tensor, logdet: Tensor, sigma: float = 1.0, elementwise_mean: bool = True) -> Tensor:
    """``WaveGlowLoss.forward`` (``model/loss.py:10-15``)."""
    loss = 0.5 * z.pow(2).sum(1) / (sigma ** 2) - logdet
    loss = loss.mean()
    if elementwise_mean:
        loss = loss / z.size(1)
    return loss


def waveglow_infer(sd: State, spec: WaveGlowSpec, h: Tensor, z: Tensor) -> Tensor:
    """``FlowBase.infer`` (``model/base.py:42-55``) with the noise ``z`` supplied by the caller
    (already scaled by sigma) so that oracle and product consume identical samples."""
    x, _ = waveglow_reverse(sd, spec, z, h)
    return x.squeeze()


# --------------------------------------------------------------------------------------------
# gradients (autograd over the naive restatement == the reference's reversible backward)
# --------------------------------------------------------------------------------------------
def _leafify(sd: State) -> State:
    return {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}


def waveglow_train_step(sd: State, spec: WaveGlowSpec, x: Tensor, h: Tensor, sigma: float
                        ) -> Tuple[Tensor, Tensor, Tensor, State]:
    """One fwd + loss + bwd; returns (z, logdet, loss, grads keyed like the state dict)."""
    leaf = _leafify(sd)
    z, logdet = waveglow_forward(leaf, spec, x, h)
    loss = waveglow_loss(z, logdet, sigma)
    keys = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return z.detach(), logdet.detach(), loss.detach(), {k: g for k, g in zip(keys, grads) if g is not None}


def coupling_grads(sd: State, prefix: str, x: Tensor, y: Tensor, dz: Tensor, dlog_s: Tensor,
                   reverse: bool = False, need_dy: bool = False):
    """Gradient of <z,dz> + <log_s,dlog_s> w.r.t. (x, params[, y]) for the coupling block in the
    forward (``AffineCouplingFunc``) or reverse (``InvAffineCouplingFunc``) direction."""
    leaf = _leafify({k: v for k, v in sd.items() if k.startswith(prefix)})
    xx = x.detach().clone().requires_grad_(True)
    yy = y.detach().clone().requires_grad_(need_dy)
    fn = coupling_reverse if reverse else coupling_forward
    out, ls = fn(leaf, prefix, xx, yy)
    obj = (out * dz).sum() + (ls * dlog_s).sum()
    keys = list(leaf.keys())
    wrt = [xx] + [leaf[k] for k in keys] + ([yy] if need_dy else [])
    g = torch.autograd.grad(obj, wrt)
    dx = g[0]
    dparams = {k: gi for k, gi in zip(keys, g[1:1 + len(keys)])}
    dy = g[-1] if need_dy else None
    return out.detach(), ls.detach(), dx, dparams, dy


def random_state(spec: WaveGlowSpec, wn_channels: int, depth: int, seed: int = 0,
                 radix: int = 3, end_std: Optional[float] = None, dtype=torch.float32) -> State:
    """Random-init weights with the reference's key layout and init distributions
    (Conv1d default init = kaiming_uniform(a=sqrt(5)); weight_norm sets g = ||v||; the 1x1 conv is
    a QR orthogonal matrix with positive determinant, ``model/efficient_modules.py:22-26``;
    ``end`` keeps the default Conv1d init when ``zero_init=False``, ``model/waveglow.py:92-96``).
    Used where the reference cannot be imported (GPU box); the distributions matter only for
    conditioning of the problem, not for parity."""
    gen = torch.Generator().manual_seed(seed)

    def conv_v(out_c, in_c, k):
        bound = 1.0 / (in_c * k) ** 0.5
        return (torch.rand(out_c, in_c, k, generator=gen, dtype=dtype) * 2 - 1) * bound

    sd: State = {}
    sd["upsampler.bias"] = (torch.rand(spec.n_mels, generator=gen, dtype=dtype) * 2 - 1) / spec.sub_win ** 0.5
    v = conv_v(spec.n_mels, 1, spec.sub_win)
    sd["upsampler.weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1)
    sd["upsampler.weight_v"] = v
    for k, c in enumerate(spec.channels):
        q = torch.linalg.qr(torch.randn(c, c, generator=gen, dtype=torch.float64))[0]
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        sd[f"invconv1x1.{k}.weight"] = q.to(dtype).contiguous().unsqueeze(-1)
    for k, c in enumerate(spec.channels):
        p = f"WNs.{k}.F."

        def put(name, out_c, in_c, ks):
            vv = conv_v(out_c, in_c, ks)
            sd[p + name + ".weight_g"] = vv.flatten(1).norm(dim=1).view(-1, 1, 1)
            sd[p + name + ".weight_v"] = vv

        put("V", 2 * wn_channels * depth, spec.n_mels, 1)
        put("start", wn_channels, c // 2, 1)
        for i in range(depth):
            put(f"layers.{i}.W", 2 * wn_channels, wn_channels, radix)
            put(f"layers.{i}.W_o", wn_channels * (2 if i < depth - 1 else 1), wn_channels, 1)
        if end_std is None:
            sd[p + "end.weight"] = conv_v(c, wn_channels, 1)
        else:
            sd[p + "end.weight"] = torch.randn(c, wn_channels, 1, generator=gen, dtype=dtype) * end_std
    return sd


# --------------------------------------------------------------------------------------------
# WSRGlow (model/wsrglow.py:8-56): a WaveGlow whose conditioning is built from the low-rate signal
# --------------------------------------------------------------------------------------------
def wsrglow_spec(upsample_rate: int = 2) -> WaveGlowSpec:
    """``WSRGlow.__init__`` (``model/wsrglow.py:21-26``): WaveGlow(12 flows, n_group 8r, early 4/2,
    hop 8r, n_mels 8*400 + 51*9 = 3659)."""
    return WaveGlowSpec(12, 8 * upsample_rate, 4, 2, 8 * upsample_rate, 8 * 400 + 51 * 9)


def mu_law_encode(x: Tensor, channels: int = 256) -> Tensor:
    """torchaudio ``MuLawEncoding(256)`` as used at ``model/wsrglow.py:27-30``:
    sign(x) log1p(mu |x|) / log1p(mu), mapped to integer codes 0..mu."""
    mu = torch.tensor(channels - 1.0, dtype=x.dtype)
    x_mu = torch.sign(x) * torch.log1p(mu * torch.abs(x)) / torch.log1p(mu)
    return ((x_mu + 1) / 2 * mu + 0.5).to(torch.int64)


def wsrglow_cond(sd: State, c: Tensor) -> Tensor:
    """``WSRGlow._get_cond`` (``model/wsrglow.py:37-50``) on a CLONE of ``c`` (the reference clips in place):
    mu-law code embedding reshaped to (B, 3200, L/8), 9 STFT magnitudes (n_fft 16, hop 8, hann, reflect
    pad 4) and 9 x 50 phase-embedding rows -> (B, 3659, L/8)."""
    c = c.clone().clip_(-1, 1)
    c_emb = F.embedding(mu_law_encode(c), sd["mu_enc.1.weight"]).view(c.shape[0], -1, 8 * 400).transpose(1, 2)
    spec = torch.stft(F.pad(c.unsqueeze(1), (4, 4), mode="reflect").squeeze(1), n_fft=16, hop_length=8,
                      window=sd["window"], center=False, return_complex=True)
    mag = spec.abs()
    emb_w = sd["angle_embed.embed.weight"]
    n_emb = emb_w.shape[0]
    index = ((spec.angle() / torch.pi + 1) * 0.5 * (n_emb - 1)).long()     # AngleEmbedding.forward, :15-18
    phase_emb = F.embedding(index, emb_w).permute(0, 1, 3, 2).reshape(spec.shape[0], 50 * 9, -1)
    return torch.cat([c_emb, mag, phase_emb], dim=1)


def wsrglow_forward(sd: State, spec: WaveGlowSpec, x: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
    """``WSRGlow.forward_computation`` (``model/wsrglow.py:52-53``)."""
    return waveglow_forward(sd, spec, x, wsrglow_cond(sd, c))


def wsrglow_reverse(sd: State, spec: WaveGlowSpec, z: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
    """``WSRGlow.reverse_computation`` (``model/wsrglow.py:55-56``)."""
    return waveglow_reverse(sd, spec, z, wsrglow_cond(sd, c))


def wsrglow_train_step(sd: State, spec: WaveGlowSpec, x: Tensor, c: Tensor, sigma: float):
    """fwd + loss + bwd of WSRGlow; gradients include both embedding tables."""
    leaf = _leafify(sd)
    leaf["window"] = sd["window"]
    z, logdet = wsrglow_forward(leaf, spec, x, c)
    loss = waveglow_loss(z, logdet, sigma)
    keys = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return z.detach(), logdet.detach(), loss.detach(), {k: g for k, g in zip(keys, grads) if g is not None}


def wsrglow_random_state(upsample_rate: int, wn_channels: int, depth: int, seed: int = 0,
                         end_std: Optional[float] = None) -> State:
    """``random_state`` plus WSRGlow's extra entries (``model/wsrglow.py:27-35``): mu-law embedding (256, 400),
    angle embedding (120, 50) -- nn.Embedding default init N(0, 1) -- and the hann(16) window buffer."""
    spec = wsrglow_spec(upsample_rate)
    sd = random_state(spec, wn_channels, depth, seed=seed, end_std=end_std)
    gen = torch.Generator().manual_seed(seed + 7919)
    sd["mu_enc.1.weight"] = torch.randn(256, 400, generator=gen)
    sd["angle_embed.embed.weight"] = torch.randn(120, 50, generator=gen)
    sd["window"] = torch.hann_window(16)
    return sd


# --------------------------------------------------------------------------------------------
# WaveFlow (model/waveflow.py:14-265): 2-D WN over the squeezed (height, width) image, row-autoregressive
# --------------------------------------------------------------------------------------------
# height dilations per n_group (``model/waveflow.py:81-87``); width dilations are 2**i (``:90-91``)
WAVEFLOW_H_DILATIONS = {
    8: [1] * 8,
    16: [1] * 8,
    32: [1, 2, 4] * 2 + [1, 2],
    64: [1, 2, 4, 8, 16, 1, 2, 4],
    128: [1, 2, 4, 8, 16, 32, 64, 1],
}


class WaveFlowSpec:
    """Structural hyper-parameters of ``WaveFlow.__init__`` (``model/waveflow.py:154-189``); hop length is
    fixed at 256 (``:163``)."""

    def __init__(self, flows: int, n_group: int, n_mels: int, use_conv1x1: bool = False):
        self.flows = flows
        self.n_group = n_group
        self.n_mels = n_mels
        self.use_conv1x1 = use_conv1x1
        self.hop = 256
        self.sub_sr = self.hop // n_group
        self.h_dilations = WAVEFLOW_H_DILATIONS[n_group]
        self.depth = 8


def waveflow_upsample_h(sd: State, spec: WaveFlowSpec, h: Tensor) -> Tensor:
    """``model/waveflow.py:169-175,263-265``: ReplicationPad1d((0, 1)) -> weight-normed DENSE
    ConvTranspose1d(n_mels, n_mels, 2*sub_sr+1, stride sub_sr, padding sub_sr//2) -> LeakyReLU(0.4).
    ``nn.utils.weight_norm`` with dim=0 normalises the (in, out, k) transposed-conv weight per INPUT channel."""
    w = resolve_weight(sd, "upsampler.1.")
    hp = F.pad(h, (0, 1), mode="replicate")
    y = F.conv_transpose1d(hp, w, sd.get("upsampler.1.bias"), stride=spec.sub_sr, padding=spec.sub_sr // 2)
    return F.leaky_relu(y, 0.4)


def wn2d_layer(sd: State, prefix: str, i: int, h_dil: int, x: Tensor, v: Tensor, last: bool):
    """``NonCausalLayer2D.forward`` (``model/waveflow.py:41-51``): causal padding 2*h_dil rows on top,
    'same' padding 2**i columns left and right, 3x3 conv with dilation (h_dil, 2**i)."""
    w = resolve_weight(sd, prefix + f"layers.{i}.W.")
    radix = w.shape[-1]
    d = 2 ** i
    pad = d * (radix - 1) // 2
    tmp = F.pad(x, [pad, pad, h_dil * (radix - 1), 0])
    xy = F.conv2d(tmp, w, resolve_bias(sd, prefix + f"layers.{i}.W."), dilation=(h_dil, d)) + v
    zw, zf = xy.chunk(2, 1)
    g = fused_gate(zw, zf)
    ro = F.conv2d(g, resolve_weight(sd, prefix + f"layers.{i}.W_o."), resolve_bias(sd, prefix + f"layers.{i}.W_o."))
    if last:
        return None, ro
    res_ch = x.shape[1]
    return ro[:, :res_ch] + x, ro[:, res_ch:]


def wn2d_forward(sd: State, prefix: str, x: Tensor, y: Tensor, h_dilations: Sequence[int]
                 ) -> Tuple[Tensor, Tensor]:
    """``WN2D.forward`` (``model/waveflow.py:128-135``).  x: (B, 1, H, W), y: (B, aux, W) -> (log_s, t)
    each (B, 1, H, W)."""
    depth = len(h_dilations)
    h = F.conv2d(x, resolve_weight(sd, prefix + "start."), resolve_bias(sd, prefix + "start."))
    v_all = F.conv1d(y, resolve_weight(sd, prefix + "V."), resolve_bias(sd, prefix + "V.")).unsqueeze(2)
    cum_skip = None
    for i, v in enumerate(v_all.chunk(depth, 1)):
        h, skip = wn2d_layer(sd, prefix, i, h_dilations[i], h, v, last=(i == depth - 1))
        cum_skip = skip if cum_skip is None else cum_skip + skip
    out = F.conv2d(cum_skip, sd[prefix + "end.weight"], sd.get(prefix + "end.bias"))
    log_s, t = out.chunk(2, 1)
    return log_s, t


def waveflow_squeeze(x: Tensor, n_group: int) -> Tensor:
    """``model/waveflow.py:195``: (B, T) -> (B, 1, n_group, T/n_group), image[h, w] = x[w*n_group + h]."""
    return x.view(x.size(0), 1, -1, n_group).transpose(2, 3).contiguous()


def waveflow_unsqueeze(x: Tensor) -> Tensor:
    """``model/waveflow.py:219,260``."""
    return x.squeeze(1).transpose(1, 2).contiguous().view(x.size(0), -1)


def waveflow_forward(sd: State, spec: WaveFlowSpec, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``WaveFlow.forward_computation`` (``model/waveflow.py:191-219``)."""
    y = waveflow_upsample_h(sd, spec, h)
    x = waveflow_squeeze(x, spec.n_group)
    y = y[..., :x.size(-1)]
    logdet = 0
    for k in range(spec.flows):
        x0 = x[:, :, :1]
        log_s, t = wn2d_forward(sd, f"WNs.{k}.", x[:, :, :-1], y, spec.h_dilations)
        xout = x[:, :, 1:] * log_s.exp() + t
        logdet = logdet + log_s.sum((1, 2, 3))
        if not spec.use_conv1x1:
            x = torch.cat((xout.flip(2), x0), 2)
        else:
            x, ldw = conv1x1_forward(sd[f"invconv1x1.{k}.weight"], torch.cat((x0, xout), 2).squeeze(1))
            x = x.unsqueeze(1)
            logdet = logdet + ldw
    return waveflow_unsqueeze(x), logdet


def wn2d_reverse_step(sd: State, prefix: str, xrow: Tensor, cond: Sequence[Tensor], buffers, h_dilations):
    """``WN2D.reverse_mode_forward`` (``model/waveflow.py:137-151``) with
    ``NonCausalLayer2D.reverse_mode_forward`` (``:53-67``): one new row ``xrow`` (B, 1, 1, W) goes through the
    eight layers; every layer keeps a rolling buffer of its last 2*h_dil+1 input rows (zeros before row 0)."""
    depth = len(h_dilations)
    x = F.conv2d(xrow, resolve_weight(sd, prefix + "start."), resolve_bias(sd, prefix + "start."))
    new_buffers = []
    cum_skip = None
    for i in range(depth):
        hd = h_dilations[i]
        w = resolve_weight(sd, prefix + f"layers.{i}.W.")
        radix = w.shape[-1]
        d = 2 ** i
        pad = d * (radix - 1) // 2
        if buffers is None:
            buf = F.pad(x, [0, 0, hd * (radix - 1), 0])
        else:
            buf = torch.cat((buffers[i][:, :, 1:], x), 2)
        new_buffers.append(buf)
        xy = F.conv2d(F.pad(buf, [pad, pad]), w, resolve_bias(sd, prefix + f"layers.{i}.W."),
                      dilation=(hd, d)) + cond[i]
        zw, zf = xy.chunk(2, 1)
        g = fused_gate(zw, zf)
        ro = F.conv2d(g, resolve_weight(sd, prefix + f"layers.{i}.W_o."),
                      resolve_bias(sd, prefix + f"layers.{i}.W_o."))
        if i < depth - 1:
            res_ch = x.shape[1]
            x, skip = ro[:, :res_ch] + x, ro[:, res_ch:]
        else:
            skip = ro
        cum_skip = skip if cum_skip is None else cum_skip + skip
    out = F.conv2d(cum_skip, sd[prefix + "end.weight"], sd.get(prefix + "end.bias"))
    log_s, t = out.chunk(2, 1)
    return log_s, t, new_buffers


def waveflow_reverse(sd: State, spec: WaveFlowSpec, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``WaveFlow.reverse_computation`` (``model/waveflow.py:221-261``): flows in reverse order; inside a flow the
    rows are generated one after the other, row i from rows < i."""
    y = waveflow_upsample_h(sd, spec, h)
    z = waveflow_squeeze(z, spec.n_group)
    y = y[..., :z.size(-1)]
    logdet = None
    for k in range(spec.flows - 1, -1, -1):
        prefix = f"WNs.{k}."
        if not spec.use_conv1x1:
            z = z.flip(2)
        else:
            z, ldw = conv1x1_reverse(sd[f"invconv1x1.{k}.weight"], z.squeeze(1))
            z = z.unsqueeze(1)
            logdet = ldw.repeat(z.shape[0]) if logdet is None else logdet + ldw
        cond = F.conv1d(y, resolve_weight(sd, prefix + "V."), resolve_bias(sd, prefix + "V.")
                        ).unsqueeze(2).chunk(spec.depth, 1)
        xnew = z[:, :, :1]
        rows = [xnew]
        buffers = None
        for i in range(1, spec.n_group):
            log_s, t, buffers = wn2d_reverse_step(sd, prefix, xnew, cond, buffers, spec.h_dilations)
            xnew = (z[:, :, i:i + 1] - t) / log_s.exp()
            rows.append(xnew)
            term = -log_s.sum((1, 2, 3))
            logdet = term if logdet is None else logdet + term
        z = torch.cat(rows, 2)
    return waveflow_unsqueeze(z), logdet


def waveflow_train_step(sd: State, spec: WaveFlowSpec, x: Tensor, h: Tensor, sigma: float):
    """fwd + loss + bwd (plain autograd, as the reference trains WaveFlow: ``memory_efficient=false``,
    ``configs/waveflow_LJ_speech.json:9-10``)."""
    leaf = _leafify(sd)
    z, logdet = waveflow_forward(leaf, spec, x, h)
    loss = waveglow_loss(z, logdet, sigma)
    keys = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return z.detach(), logdet.detach(), loss.detach(), {k: g for k, g in zip(keys, grads) if g is not None}


def waveflow_random_state(spec: WaveFlowSpec, wn_channels: int, seed: int = 0,
                          end_std: Optional[float] = None, dtype=torch.float32) -> State:
    """Random-init weights with the reference's WaveFlow key layout (``upsampler.1.*``, ``WNs.k.*`` without the
    ``.F`` level, 4-D Conv2d weights) and default-init distributions; see ``random_state``."""
    gen = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        return (torch.rand(*shape, generator=gen, dtype=dtype) * 2 - 1) / fan_in ** 0.5

    def gv(v):
        return v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))

    sd: State = {}
    K = 2 * spec.sub_sr + 1
    v = uni((spec.n_mels, spec.n_mels, K), spec.n_mels * K)
    sd["upsampler.1.bias"] = uni((spec.n_mels,), spec.n_mels * K)
    sd["upsampler.1.weight_g"] = gv(v)
    sd["upsampler.1.weight_v"] = v
    C = wn_channels
    for k in range(spec.flows):
        if spec.use_conv1x1:
            q = torch.linalg.qr(torch.randn(spec.n_group, spec.n_group, generator=gen, dtype=torch.float64))[0]
            if torch.det(q) < 0:
                q[:, 0] = -q[:, 0]
            sd[f"invconv1x1.{k}.weight"] = q.to(dtype).contiguous().unsqueeze(-1)
    for k in range(spec.flows):
        p = f"WNs.{k}."

        def put(name, shape, fan_in):
            vv = uni(shape, fan_in)
            sd[p + name + ".weight_g"] = gv(vv)
            sd[p + name + ".weight_v"] = vv

        put("V", (2 * C * spec.depth, spec.n_mels, 1), spec.n_mels)
        put("start", (C, 1, 1, 1), 1)
        for i in range(spec.depth):
            put(f"layers.{i}.W", (2 * C, C, 3, 3), C * 9)
            put(f"layers.{i}.W_o", (C * (2 if i < spec.depth - 1 else 1), C, 1, 1), C)
        if end_std is None:
            sd[p + "end.weight"] = uni((2, C, 1, 1), C)
        else:
            sd[p + "end.weight"] = torch.randn(2, C, 1, 1, generator=gen, dtype=dtype) * end_std
    return sd


# --------------------------------------------------------------------------------------------
# MRWaveGlow (model/mr_waveglow.py): Haar-like band split over channels, per-level flows, prior flows
# --------------------------------------------------------------------------------------------
@dataclass
class MRSpec:
    prior_flows: int
    n_group: int
    hop_size: int
    n_mels: int
    levels: int = 3
    flows: int = 4
    super_resolution: bool = False

    @property
    def upsample_factor(self) -> int:
        return self.hop_size // self.n_group


def mr_upsample_h(spec: MRSpec, h: Tensor) -> Tensor:
    """``model/mr_waveglow.py:133-134``: linear interpolation of the mel to the squeezed rate."""
    return F.interpolate(h, scale_factor=spec.upsample_factor, mode="linear")


def _mr_steps(sd: State, conv_prefix: str, wn_prefix: str, n: int, x: Tensor, cond: Tensor, inverse: bool):
    total = 0
    for k in (range(n - 1, -1, -1) if inverse else range(n)):
        w = sd[f"{conv_prefix}{k}.weight"]
        if inverse:
            x, log_s = coupling_reverse(sd, f"{wn_prefix}{k}.F.", x, cond)
            x, ldw = conv1x1_reverse(w, x)
        else:
            x, ldw = conv1x1_forward(w, x)
            x, log_s = coupling_forward(sd, f"{wn_prefix}{k}.F.", x, cond)
        total = total + ldw + log_s.sum((1, 2))
    return x, total


def mrwaveglow_forward(sd: State, spec: MRSpec, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``MRWaveGlow.forward_computation`` (``model/mr_waveglow.py:60-92``)."""
    y = mr_upsample_h(spec, h)
    B = x.size(0)
    x = x.view(B, -1, spec.n_group).transpose(1, 2)
    assert x.size(2) <= y.size(2)
    y = y[..., :x.size(2)]
    outs, logdet = [], 0
    for level in range(spec.levels - 1):
        x0, x1 = x[:, ::2], x[:, 1::2]
        x_diff, x = x1 - x0, (x0 + x1) * 0.5
        cond = x if spec.super_resolution else torch.cat([x, y], 1)
        x_diff, ld = _mr_steps(sd, f"invconv1x1_list.{level}.", f"WNs_list.{level}.", spec.flows, x_diff, cond, False)
        logdet = logdet + ld
        outs.append(x_diff)
    x, ld = _mr_steps(sd, "prior_invconv1x1.", "prior_WNs.", spec.prior_flows, x, y, False)
    outs.append(x)
    return torch.cat(outs, 1).transpose(1, 2).contiguous().view(B, -1), logdet + ld


def mrwaveglow_reverse(sd: State, spec: MRSpec, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``MRWaveGlow.reverse_computation`` (``model/mr_waveglow.py:94-131``)."""
    y = mr_upsample_h(spec, h)
    B = z.size(0)
    z = z.view(B, -1, spec.n_group).transpose(1, 2)
    assert z.size(2) <= y.size(2)
    y = y[..., :z.size(2)]
    remained = []
    for _ in range(spec.levels - 1):
        r, z = z.chunk(2, 1)
        remained.append(r)
    z, logdet = _mr_steps(sd, "prior_invconv1x1.", "prior_WNs.", spec.prior_flows, z, y, True)
    for level in range(spec.levels - 2, -1, -1):
        z_diff = remained.pop()
        cond = z if spec.super_resolution else torch.cat([z, y], 1)
        z_diff, ld = _mr_steps(sd, f"invconv1x1_list.{level}.", f"WNs_list.{level}.", spec.flows, z_diff, cond, True)
        logdet = logdet + ld
        z_0, z_1 = z - z_diff * 0.5, z + z_diff * 0.5
        z = torch.stack([z_0, z_1], 2).view(B, -1, z_0.size(2))
    return z.transpose(1, 2).contiguous().view(B, -1), logdet


def mrwaveglow_train_step(sd: State, spec: MRSpec, x: Tensor, h: Tensor, sigma: float):
    """One fwd + loss + bwd; returns (z, logdet, loss, grads keyed like the state dict)."""
    leaf = _leafify(sd)
    z, logdet = mrwaveglow_forward(leaf, spec, x, h)
    loss = waveglow_loss(z, logdet, sigma)
    keys = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return z.detach(), logdet.detach(), loss.detach(), {k: g for k, g in zip(keys, grads) if g is not None}


# --------------------------------------------------------------------------------------------
# MelGlow (model/melglow.py): kernel predictor, location-variable convolution layers, flow wiring
# --------------------------------------------------------------------------------------------
@dataclass
class MelGlowSpec:
    flows: int
    n_group: int
    n_early_every: int
    n_early_size: int
    hop_size: int
    n_mels: int
    # WN_LVC
    dilation_channels: int = 48
    residual_channels: int = 48
    skip_channels: int = 48
    depth: int = 7
    radix: int = 3
    predict_channels: int = 64
    predict_layers: int = 3

    @property
    def upsample_factor(self) -> int:
        return self.hop_size // self.n_group


def _bn_train(sd: State, prefix: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """``nn.BatchNorm1d`` in training mode: batch statistics over (batch, time), biased variance."""
    mean = x.mean((0, 2), keepdim=True)
    var = x.var((0, 2), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * sd[prefix + "weight"].view(1, -1, 1) + sd[prefix + "bias"].view(1, -1, 1)


def lvc_predictor(sd: State, prefix: str, spec: MelGlowSpec, y: Tensor) -> Tensor:
    """``Predictor.forward`` (``model/melglow.py:44-49``), BatchNorm in training mode."""
    g = spec.depth
    s = torch.tanh(_bn_train(sd, prefix + "start.1.", F.conv1d(y, sd[prefix + "start.0.weight"], sd.get(prefix + "start.0.bias"))))
    for j in range(spec.predict_layers):
        b = prefix + f"res_blocks.{j}."
        u = torch.tanh(_bn_train(sd, b + "1.", F.conv1d(s, sd[b + "0.weight"], sd.get(b + "0.bias"), groups=g)))
        u = torch.tanh(_bn_train(sd, b + "4.", F.conv1d(u, sd[b + "3.weight"], sd.get(b + "3.bias"), groups=g)))
        s = u + s
    return F.conv1d(s, sd[prefix + "end.weight"], sd.get(prefix + "end.bias"), groups=g)


def lvc_layer(sd: State, prefix: str, dilation: int, radix: int, x: Tensor, weights: Tensor, last: bool):
    """``NonCausalLayerLVC.forward`` (``model/melglow.py:74-92``): unfold the padded signal into one window per frame and
    apply that frame's kernel as a grouped convolution."""
    batch, steps, cout, cin, _ = weights.shape
    pad = dilation * (radix - 1) // 2
    offset = x.shape[2] // steps
    w = weights.reshape(batch * steps * cout, cin, radix)
    ux = F.pad(x, (pad, pad)).unfold(2, pad * 2 + offset, offset).transpose(1, 2).contiguous().view(1, -1, pad * 2 + offset)
    z = F.conv1d(ux, w, dilation=dilation, groups=batch * steps)
    zw, zv = z.view(batch, steps, cout, -1).transpose(1, 2).contiguous().view(batch, cout, -1).chunk(2, 1)
    o = F.conv1d(fused_gate(zw, zv), resolve_weight(sd, prefix + "W_o."), resolve_bias(sd, prefix + "W_o."))
    if last:
        return None, o
    cr = x.shape[1]
    return o[:, :cr] + x, o[:, cr:]


def wn_lvc_forward(sd: State, prefix: str, spec: MelGlowSpec, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """``WN_LVC.forward`` (``model/melglow.py:147-159``)."""
    x = F.conv1d(x, resolve_weight(sd, prefix + "start."), resolve_bias(sd, prefix + "start."))
    wts = lvc_predictor(sd, prefix + "pred.", spec, y)
    wts = wts.view(wts.shape[0], spec.depth, -1, wts.shape[2]).permute(1, 0, 3, 2).contiguous()
    cum = 0
    for i in range(spec.depth):
        w = wts[i].view(wts.shape[1], wts.shape[2], 2 * spec.dilation_channels, spec.residual_channels, spec.radix)
        x, skip = lvc_layer(sd, prefix + f"layers.{i}.", 2 ** i, spec.radix, x, w, i == spec.depth - 1)
        cum = cum + skip
    out = F.conv1d(cum, sd[prefix + "end.weight"], sd.get(prefix + "end.bias"))
    return tuple(out.chunk(2, 1))


def _lvc_coupling(sd: State, prefix: str, spec: MelGlowSpec, x: Tensor, y: Tensor, inverse: bool):
    """``AffineCouplingBlock`` around ``WN_LVC`` (``model/efficient_modules.py:77-82,91-96``)."""
    xa, xb = x.chunk(2, 1)
    log_s, t = wn_lvc_forward(sd, prefix, spec, xa, y)
    if inverse:
        return torch.cat((xa, (xb - t) / log_s.exp()), 1), -log_s
    return torch.cat((xa, xb * log_s.exp() + t), 1), log_s


def melglow_forward(sd: State, spec: MelGlowSpec, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``MelGlow.forward_computation`` (``model/melglow.py:204-232``)."""
    B = x.size(0)
    x = x[:, :x.shape[1] // spec.hop_size * spec.hop_size]
    x = x.view(B, -1, spec.n_group).transpose(1, 2)
    y = h[..., :x.shape[2] // spec.upsample_factor]
    outs, logdet = [], 0
    for k in range(spec.flows):
        if k % spec.n_early_every == 0 and k:
            outs.append(x[:, :spec.n_early_size])
            x = x[:, spec.n_early_size:]
        x, ldw = conv1x1_forward(sd[f"invconv1x1.{k}.weight"], x)
        x, log_s = _lvc_coupling(sd, f"WNs.{k}.F.", spec, x, y, False)
        logdet = logdet + ldw + log_s.sum((1, 2))
    outs.append(x)
    return torch.cat([o.transpose(1, 2) for o in outs], 2).reshape(B, -1), logdet


def melglow_reverse(sd: State, spec: MelGlowSpec, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """``MelGlow.reverse_computation`` (``model/melglow.py:234-258``)."""
    B = z.size(0)
    z = z[:, :z.shape[1] // spec.hop_size * spec.hop_size]
    z = z.view(B, -1, spec.n_group).transpose(1, 2)
    y = h[..., :z.shape[2] // spec.upsample_factor]
    sizes, rem = [], spec.n_group
    for k in range(spec.flows):
        if k % spec.n_early_every == 0 and k:
            rem -= spec.n_early_size
            sizes.append(spec.n_early_size)
    sizes.append(rem)
    *remained, z = z.split(sizes, 1)
    remained = list(remained)
    logdet = 0
    for k in range(spec.flows - 1, -1, -1):
        z, log_s = _lvc_coupling(sd, f"WNs.{k}.F.", spec, z, y, True)
        z, ldw = conv1x1_reverse(sd[f"invconv1x1.{k}.weight"], z)
        logdet = logdet + ldw + log_s.sum((1, 2))
        if k % spec.n_early_every == 0 and k:
            z = torch.cat((remained.pop(), z), 1)
    return z.transpose(1, 2).contiguous().view(B, -1), logdet


def melglow_train_step(sd: State, spec: MelGlowSpec, x: Tensor, h: Tensor, sigma: float):
    leaf = _leafify(sd)
    z, logdet = melglow_forward(leaf, spec, x, h)
    loss = waveglow_loss(z, logdet, sigma)
    keys = [k for k, v in leaf.items() if v.requires_grad and "running_" not in k]
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return z.detach(), logdet.detach(), loss.detach(), {k: g for k, g in zip(keys, grads) if g is not None}
