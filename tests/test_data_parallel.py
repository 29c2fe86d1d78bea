"""Host-side data-parallel logic on CPU: world_size-2 gloo processes.

The compute kernels need a GPU; what is covered here is everything around them that changes with N > 1: the
bucket table over the flat gradient buffer, the bucketed sum-all-reduce issued in backward-completion order and
its 1/world scale, replica consistency of a (reference) SGD update, loss-vector reduction, batch sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aldi_b200 import data_parallel as dp
from aldi_b200.detector import FlatLayout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_bucket_ranges_cover_trainable_range():
    layout = FlatLayout(8)
    r = dp.bucket_ranges(layout)
    assert list(r) == ["res3", "res4", "res5", "fpn", "heads"]          # forward (layout) order
    pos = 0
    for tag, (a, b) in r.items():
        assert a == pos and b > a, (tag, a, b)
        pos = b
    assert pos == layout.num_trainable
    # every trainable tensor lies inside exactly the bucket of its stage
    for (layer, field), (off, n, key, _) in layout.entries.items():
        if off >= layout.num_trainable:
            continue
        tag = layer[:4] if layer.startswith("res") else "fpn" if layer.startswith("fpn") else "heads"
        a, b = r[tag]
        assert a <= off and off + n <= b, key
    # frozen stem / res2 and the FrozenBN buffers are outside the reduced range
    assert layout.entries[("stem", "weight")][0] >= layout.num_trainable
    assert layout.entries[("res2.0.conv1", "weight")][0] >= layout.num_trainable


def test_single_process_reducer_is_a_noop():
    layout = FlatLayout(8)
    g = torch.ones(layout.num_trainable)
    red = dp.GradReducer(layout, g, None)
    assert not red.active
    red.ready("heads")
    assert red.finish() == 1.0 and red.collectives == 0
    assert torch.equal(g, torch.ones_like(g))


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = FlatLayout(8)
        nt = layout.num_trainable
        gen = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(nt, generator=gen)
        local = grad.clone()
        red = dp.GradReducer(layout, grad, dist.group.WORLD)
        assert red.active and red.world == world
        # the last backward reports buckets in completion order; "res3" is deliberately left to finish()
        for tag in ("heads", "fpn", "res5", "res4"):
            red.ready(tag)
            red.ready(tag)  # idempotent within a step
        scale = red.finish()
        assert scale == 1.0 / world and red.collectives == 5
        # reference: one all-reduce of the whole buffer
        ref = local.clone()
        dist.all_reduce(ref)
        assert torch.equal(grad, ref), "bucketed all-reduce != whole-buffer all-reduce"
        # a second step reuses the reducer
        grad.copy_(local)
        assert red.finish() == 1.0 / world and torch.equal(grad, ref)
        # replicas stay identical after the (reference) SGD update with the folded 1/world scale
        p = torch.linspace(-1, 1, nt)
        m = torch.zeros(nt)
        g = grad * scale + 1e-4 * p
        m = 0.9 * m + g
        p = p - 0.01 * m
        assert dp.check_replicas_in_sync(p, dist.group.WORLD) == 0.0
        # loss vector for logging: mean over ranks
        lv = dp.reduce_loss_vector(torch.tensor([1.0 + rank, 10.0 * (rank + 1)]), dist.group.WORLD)
        assert torch.allclose(lv, torch.tensor([1.5, 15.0]))
        # global batch sharding
        shard = dp.shard_for_rank(list(range(8)), rank, world)
        assert shard == list(range(rank * 4, rank * 4 + 4))
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bucketed_allreduce_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
