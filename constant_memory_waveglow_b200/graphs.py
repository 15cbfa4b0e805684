"""One training step as ONE CUDA graph.

A constant-memory training step of the LJ configuration is ~460 kernel launches: 36 long tensor-core kernels and a few
hundred short ones (1x1 convolutions, couplings, reductions, weight-norm backward, optimizer), each issued from Python
through autograd.  The long kernels hide the host behind them; the runs of short ones between them do not, and at the
reference's data-parallel split (``train.py:51-53``: global batch 24 // 8 GPUs = 3 segments per GPU) EVERY kernel is short
and the step is bound by the host issuing launches.  Everything in the step has static shapes and -- with the per-flow
gradient buckets of ``parallel.FlowGradSync`` and the fused optimizer -- static addresses, so the whole of

    zero grads -> forward -> loss -> reversible backward (+ per-flow NCCL all-reduces) -> average -> optimizer step

is captured once per input shape and replayed with a single launch.  The arithmetic and its order are exactly those of the
eager step (same kernels, same stream order); ``tests/test_gpu_graphs.py`` checks bit-identical parameters after several steps.

    step = GraphedTrainStep(model, lambda x, h: loss_fn(*model(x, h)), optimizer, sync)   # sync: parallel.FlowGradSync
    loss = step(x, h)                                              # tensors on the device or in pinned host memory

The returned loss is a static device tensor that the next call overwrites; read it (``.item()``) before calling again.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib as L


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_step: Callable, optimizer: torch.optim.Optimizer, sync,
                 warmup: int = 3, max_graphs: int = 2, clip_grad_norm: Optional[float] = None):
        """`loss_step(*inputs)` runs the forward pass and returns the loss, or `(loss, aux)` with `aux` a dict of tensors
        (metrics) that are kept as static outputs next to the loss (``self.aux`` after each call).  The optimizer must keep
        its state on the device (``torch.optim.Adam(..., fused=True, capturable=True)`` or any capturable optimizer)."""
        self.model, self.loss_step, self.opt, self.sync = model, loss_step, optimizer, sync
        self.warmup, self.max_graphs, self.clip = warmup, max_graphs, clip_grad_norm
        self.aux: Dict[str, torch.Tensor] = {}
        self._stream = torch.cuda.Stream(next(model.parameters()).device)
        self._graphs: Dict[Tuple, Tuple] = {}
        self._seen: Dict[Tuple, int] = {}
        self.launches_per_step = 0       # kernels of this library inside one captured step
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedTrainStep needs a capturable optimizer, e.g. Adam(..., fused=True, capturable=True)")

    # the eager step: also what is captured
    def eager(self, *inputs):
        self.sync.zero_grad()
        out = self.loss_step(*inputs)
        loss, self.aux = out if isinstance(out, tuple) else (out, {})
        loss.backward()
        self.sync.finish()
        if self.clip:
            torch.nn.utils.clip_grad_norm_([p for p in self.model.parameters() if p.requires_grad], self.clip)
        self.opt.step()
        return loss

    def _capture(self, key, inputs):
        from .waveglow import _cond_cache, invalidate_packs
        static_in = tuple(torch.empty(t.shape, dtype=t.dtype, device=self._device) for t in inputs)
        for s, t in zip(static_in, inputs):
            s.copy_(t, non_blocking=True)
        # Warm-up off the capture stream: lazy initialisation (kernel attributes, weight packs, TMA descriptors, NCCL
        # communicators, optimizer state tensors) must happen before the capture.  The warm-up steps are real steps, so the
        # parameters and the optimizer state are put back afterwards: capturing executes nothing, and the first replay is
        # then exactly ONE step from the state the caller handed over.
        params = [p for g in self.opt.param_groups for p in g["params"]]
        with torch.no_grad():
            p_keep = [p.detach().clone() for p in params]
            s_keep = {id(p): {k: v.clone() for k, v in self.opt.state[p].items() if torch.is_tensor(v)}
                      for p in params if p in self.opt.state}
        # Warm-up and capture run on ONE side stream: autograd creates each parameter's AccumulateGrad node on the stream of
        # its first use and runs it (and the post-accumulate hooks of FlowGradSync, which may copy a gradient into its
        # bucket) on THAT stream ever after -- a node born on another stream would do its work outside the capture.
        cur = torch.cuda.current_stream(self._device)
        side = self._stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(self.warmup, 1)):
                self.eager(*static_in)
        cur.wait_stream(side)
        with torch.no_grad():
            for p, k in zip(params, p_keep):
                p.copy_(k)
            for p in params:
                for name, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        old = s_keep.get(id(p), {}).get(name)
                        v.copy_(old) if old is not None else v.zero_()   # fresh Adam-family state is all zeros
        del p_keep, s_keep
        torch.cuda.synchronize(self._device)
        _cond_cache.clear()                  # conditioning slabs cached by eager calls must be re-packed INSIDE the graph
        invalidate_packs()                   # ... and so must the weight packs: the optimizer step is inside the graph too
        lib = L.load()
        before = int(lib.cmwg_launch_count())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            loss = self.eager(*static_in)
        self.launches_per_step = int(lib.cmwg_launch_count()) - before
        _cond_cache.clear()                  # ... and slabs that live in the graph's private pool must not serve eager calls
        while len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = (graph, static_in, loss, self.aux)

    def __call__(self, *inputs):
        self._device = next(self.model.parameters()).device
        key = tuple((tuple(t.shape), t.dtype) for t in inputs)
        ent = self._graphs.get(key)
        if ent is None:
            # a shape seen for the first time runs eagerly (a ragged last batch is not worth a capture); the second sight captures
            n = self._seen.get(key, 0)
            self._seen[key] = n + 1
            if n == 0 and self._graphs:
                return self.eager(*(t.to(self._device, non_blocking=True) for t in inputs))
            self._capture(key, inputs)
            ent = self._graphs[key]
        graph, static_in, loss, self.aux = ent
        for s, t in zip(static_in, inputs):
            s.copy_(t, non_blocking=True)
        graph.replay()
        return loss
