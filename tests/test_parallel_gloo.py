"""Multi-rank host logic on CPU (gloo, world_size 2): per-flow gradient buckets are reduced as soon
as each flow's backward finishes, gradients end up as the cross-rank mean, utterance sharding has no
overlap.  The CUDA kernels are not involved (tiny stand-in flows built from nn.Linear)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from constant_memory_waveglow_b200.parallel import FlowGradSync, flow_buckets, shard_utterances


class ToyFlowModel(nn.Module):
    def __init__(self, flows=3, width=5):
        super().__init__()
        self.upsampler = nn.Linear(width, width)
        self.invconv1x1 = nn.ModuleList(nn.Linear(width, width, bias=False) for _ in range(flows))
        self.WNs = nn.ModuleList(nn.Sequential(nn.Linear(width, width), nn.Tanh(), nn.Linear(width, width))
                                 for _ in range(flows))

    def forward(self, x):
        y = self.upsampler(x)
        for c, w in zip(self.invconv1x1, self.WNs):
            x = c(x)
            x = x + w(x) * y
        return x.pow(2).mean()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = ToyFlowModel()
        ref = ToyFlowModel()
        ref.load_state_dict(model.state_dict())
        xs = [torch.randn(4, 5, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        # expected: mean over ranks of the local gradients
        expect = None
        for r in range(world):
            ref.zero_grad()
            ref(xs[r]).backward()
            g = [p.grad.clone() for p in ref.parameters()]
            expect = g if expect is None else [a + b for a, b in zip(expect, g)]
        expect = [g / world for g in expect]

        buckets = flow_buckets(model)
        assert len(buckets) == 4 and len(buckets[0]) == 5       # flow 2 first, upsampler last
        sync = FlowGradSync(buckets)
        for step in range(2):
            sync.zero_grad()
            model(xs[rank]).backward()
            sync.finish()
            # buckets are launched in backward-completion order: last flow first, upsampler last
            assert sync.launch_order == [0, 1, 2, 3], sync.launch_order
            for p, e in zip(model.parameters(), expect):
                assert torch.allclose(p.grad, e, atol=1e-6), (step, rank)
        # gradients are views of the flat communication buffers (no copies)
        p0 = buckets[0][0]
        assert p0.grad.data_ptr() == sync.flat[0].data_ptr()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_flow_grad_sync_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_single_process_buckets_and_sharding():
    model = ToyFlowModel(flows=2)
    sync = FlowGradSync(flow_buckets(model))
    sync.zero_grad()
    model(torch.randn(3, 5)).backward()
    sync.finish()
    assert sync.launch_order == [0, 1, 2]
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in model.parameters())
    shards = [shard_utterances(10, r, 4) for r in range(4)]
    assert sorted(sum(shards, [])) == list(range(10))
    assert shards[0] == [0, 4, 8] and shards[3] == [3, 7]
