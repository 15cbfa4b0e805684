"""Data-parallel gradient exchange for the reversible backward (one process per GPU).

The reference gets data parallelism from Lightning's DDPPlugin (``train.py:51-53,73-78``): every rank
holds a full replica, the only exchange is one sum of fp32 gradients per step.  Here the exchange is
organised around the structure of the constant-memory backward: flows finish in the order
K-1 ... 0, and each coupling / 1x1-conv Function returns ALL gradients of its flow at once, so the
natural bucket is "one flow" (~17.9 MB fp32 at the LJ config) plus one bucket for the upsampler.

  * every bucket owns one flat fp32 buffer; the parameters' ``.grad`` tensors are VIEWS into it.  The
    fused backward kernels write their gradients DIRECTLY into those views (``grad_buffer``): with
    ``p.grad is None`` autograd adopts the returned view as ``p.grad`` without a copy or an add kernel
    (37 parameters x 12 flows would otherwise cost ~450 tiny accumulate launches per step).  Gradients
    produced by anything else (generic transforms, CPU tests) are copied into the view once;
  * a post-accumulate-grad hook counts arrivals; when a bucket is complete its all-reduce is issued
    immediately with ``async_op=True`` -- NCCL runs it on its own stream over NVLink/NVSwitch while
    the compute stream proceeds with the next flow's recompute + gradient GEMMs;
  * ``finish()`` waits for the outstanding collectives and averages (NCCL: the collective itself averages,
    ``ReduceOp.AVG``; other backends: one fused multiply over all buckets);
  * all buckets are windows of ONE flat buffer, so the exchange can also be issued as a single all-reduce at
    the end of the backward: ``mode="deferred"`` (``CMWG_GRAD_SYNC=deferred``), the DEFAULT.  The persistent
    tensor-core kernels of the WN hold every SM with one CTA each and need all their CTA pairs resident; an NCCL
    kernel that gets an SM first delays the whole grid by its own run time, once per bucket (measured: on
    8 x B200 the forward task kernel slows 0.46 -> 0.55 ms per launch under the overlapped exchange, 5 % of the
    step; on 2 x B200 10.0 -> 11.8 ms per step of forward task kernels in eager steps, and the graphed step is
    24.55 ms deferred against 24.79 ms overlapped).  One 215 MB all-reduce over NVSwitch is 0.3-0.5 ms
    unhidden.  ``mode="overlap"`` (``CMWG_GRAD_SYNC=overlap``) keeps the per-flow exchange during the backward.

Works with any process group backend (``nccl`` on the B200 box, ``gloo`` in the CPU tests).
Inference shards independent utterances across ranks and needs no collective (``shard_utterances``).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def grad_buffer(p: torch.Tensor) -> torch.Tensor:
    """Where a fused backward should write the gradient of parameter `p`: a fresh view of the flat
    communication buffer when `p` belongs to a FlowGradSync and has no gradient yet (autograd then
    adopts it as ``p.grad`` as is), else a new tensor (ordinary accumulation semantics)."""
    slot = getattr(p, "_cmwg_grad_slot", None)
    if slot is not None and p.grad is None:
        flat, off = slot
        return flat[off:off + p.numel()].view_as(p)
    return torch.empty_like(p)


def flow_buckets(model) -> List[List[torch.nn.Parameter]]:
    """One bucket per flow (1x1 conv + coupling WN), in BACKWARD completion order, then the rest."""
    buckets: List[List[torch.nn.Parameter]] = []
    seen = set()
    if hasattr(model, "WNs") and hasattr(model, "invconv1x1"):
        for k in range(len(model.WNs) - 1, -1, -1):
            ps = [p for p in list(model.WNs[k].parameters()) + list(model.invconv1x1[k].parameters())
                  if p.requires_grad]
            for p in ps:
                seen.add(id(p))
            if ps:
                buckets.append(ps)
    rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    if rest:
        buckets.append(rest)
    return buckets


class FlowGradSync:
    def __init__(self, buckets: Sequence[Sequence[torch.nn.Parameter]], process_group=None, mode: Optional[str] = None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        requested = mode or os.environ.get("CMWG_GRAD_SYNC")
        self.mode = (requested or "deferred").lower()
        if self.mode not in ("overlap", "deferred"):
            raise ValueError(f"FlowGradSync: mode must be 'overlap' or 'deferred', got {self.mode!r}")
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then one fused multiply
        self._avg_in_collective = bool(dist.is_initialized() and dist.get_backend(process_group) == "nccl")
        self.buckets = [list(b) for b in buckets]
        self.flat: List[torch.Tensor] = []
        self._pending: List[int] = []
        self._works = []
        self._launch_order: List[int] = []
        self._handles = []
        # one allocation for all buckets (each window starts on a 256-byte boundary) when they share device and dtype
        sizes = [sum(p.numel() for p in params) for params in self.buckets]
        firsts = [params[0] for params in self.buckets]
        uniform = len({(p.device, p.dtype) for p in firsts}) <= 1
        self.whole: Optional[torch.Tensor] = None
        starts: List[int] = []
        if uniform and firsts:
            off = 0
            for n in sizes:
                starts.append(off)
                off += (n + 63) // 64 * 64
            self.whole = torch.zeros(off, device=firsts[0].device, dtype=firsts[0].dtype)
        elif self.mode == "deferred":
            if requested:
                raise ValueError("FlowGradSync: mode='deferred' needs all parameters on one device with one dtype")
            self.mode = "overlap"       # mixed devices / dtypes: no single flat buffer, exchange per bucket
        for bi, params in enumerate(self.buckets):
            n = sizes[bi]
            p0 = params[0]
            flat = self.whole[starts[bi]:starts[bi] + n] if self.whole is not None else \
                torch.zeros(n, device=p0.device, dtype=p0.dtype)
            self.flat.append(flat)
            off = 0
            for p in params:
                p._cmwg_grad_slot = (flat, off)
                p.grad = None
                off += p.numel()
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
            self._pending.append(len(params))

    @staticmethod
    def _adopt(param) -> None:
        """Make `param.grad` the view of its flat-buffer slot (copying once if it was produced elsewhere)."""
        flat, off = param._cmwg_grad_slot
        view = flat[off:off + param.numel()].view_as(param)
        g = param.grad
        if g is None:
            view.zero_()
        elif g.data_ptr() != view.data_ptr() or g.stride() != view.stride():
            view.copy_(g)
        else:
            return
        param.grad = view

    def _make_hook(self, bi: int):
        def hook(param):
            self._adopt(param)
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _reduce_op(self):
        return dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM

    def _launch(self, bi: int) -> None:
        self._launch_order.append(bi)
        if self.world > 1 and self.mode == "overlap":
            self._works.append(dist.all_reduce(self.flat[bi], op=self._reduce_op(), group=self.group, async_op=True))

    def zero_grad(self) -> None:
        """Drop all gradients (``p.grad = None``) so that the next backward writes / adopts them in place;
        call before every backward."""
        self._works.clear()
        self._launch_order.clear()
        for bi, params in enumerate(self.buckets):
            for p in params:
                p.grad = None
            self._pending[bi] = len(params)

    def finish(self) -> None:
        """Wait for the collectives issued during backward and turn sums into means."""
        for bi, left in enumerate(self._pending):
            if left != 0 and left != len(self.buckets[bi]):
                raise RuntimeError(f"FlowGradSync: bucket {bi} received only part of its gradients")
            if left == len(self.buckets[bi]):
                # bucket untouched this step (unused parameters): zero gradients, still reduced so ranks stay in step
                for p in self.buckets[bi]:
                    self._adopt(p)
                if self.world > 1:
                    self._launch(bi)
        if self.world > 1 and self.mode == "deferred":
            self._works.append(dist.all_reduce(self.whole, op=self._reduce_op(), group=self.group, async_op=True))
        for w in self._works:
            w.wait()
        self._works.clear()
        if self.world > 1 and not self._avg_in_collective:
            inv = 1.0 / self.world
            if self.whole is not None:
                self.whole.mul_(inv)
            else:
                torch._foreach_mul_(self.flat, inv)

    @property
    def launch_order(self) -> List[int]:
        return list(self._launch_order)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
        for params in self.buckets:
            for p in params:
                if hasattr(p, "_cmwg_grad_slot"):
                    del p._cmwg_grad_slot


def shard_utterances(n_items: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """Round-robin utterance -> rank assignment for synthesis (no data-path collective)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return list(range(rank, n_items, world))


def allreduce_scalars(values: Iterable[float], device) -> List[float]:
    """Mean of a few python scalars across ranks (the metric sync of model/lightning.py:58-64)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
        t /= dist.get_world_size()
    return t.tolist()
