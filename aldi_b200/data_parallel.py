"""Data-parallel plumbing of the train step: one process per GPU, parameters replicated, ONE sum-all-reduce of
the flat gradient buffer per optimizer step, issued bucket by bucket while the last micro-batch's backward is
still running (NCCL over NVLink 5 / NVSwitch via torch.distributed; gloo on CPU in the tests).

Reference behaviour this replaces (aldi/dropin.py:84-85 `create_ddp_model` -> DistributedDataParallel with
`broadcast_buffers=False`, aldi/dropin.py:53): DDP mean-all-reduces the gradients inside EVERY micro-batch
backward (`aldi/trainer.py:79` calls backward once per micro-batch when SOLVER.BACKWARD_AT_END is False), i.e.
2 x `num_grad_accum_steps` collectives of the whole 165 MB gradient per step.  Averaging is linear, so
accumulating locally and reducing once gives the same gradient; the 1/world factor is folded into the fused
SGD kernel's `grad_scale`.

The flat gradient buffer is laid out in forward order (res3, res4, res5, FPN, RPN head, box head) and the
explicit backward finishes those ranges in reverse, so a bucket is a contiguous slice that becomes final at a
known point of `Detector.backward`; `GradReducer.ready(tag)` is called there.
"""
from collections import OrderedDict

import torch.distributed as dist

# backward completion order of the buckets (Detector.backward calls ready() with these tags)
BUCKET_ORDER = ("heads", "fpn", "res5", "res4", "res3")


def bucket_ranges(layout):
    """{tag: (begin, end)} over the trainable range of the flat buffer: contiguous, disjoint, covering."""
    first = {}
    for (layer, field), (off, n, _, _) in layout.entries.items():
        if off >= layout.num_trainable:
            continue
        if layer.startswith("res"):
            tag = layer[:4]
        elif layer.startswith("fpn"):
            tag = "fpn"
        else:
            tag = "heads"
        first[tag] = min(first.get(tag, off), off)
    tags = sorted(first, key=lambda t: first[t])
    out = OrderedDict()
    for i, t in enumerate(tags):
        end = first[tags[i + 1]] if i + 1 < len(tags) else layout.num_trainable
        out[t] = (first[t], end)
    begin = min(b for b, _ in out.values())
    assert begin == 0, "trainable range must start at offset 0"
    return out


class GradReducer:
    """Issues the per-step gradient all-reduce in buckets; `finish()` returns the scale (1/world) the optimizer
    kernel applies.  With no process group (single GPU) every method is a no-op and the scale is 1."""

    def __init__(self, layout, grad, process_group=None):
        self.grad, self.pg = grad, process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.ranges = bucket_ranges(layout)
        assert set(self.ranges) <= set(BUCKET_ORDER), sorted(self.ranges)
        self.pending = []
        self.done = set()
        self.collectives = 0
        if self.active:
            self.warm_up()

    def warm_up(self):
        """Run every bucket's all-reduce once on a scratch buffer BEFORE the first kernel of the step exists: NCCL sets up
        its per-algorithm connections lazily at the first collective of a given size class (runtime connect: cuMem
        allocations, IPC / multicast mappings, NVLS buffers on 8 GPUs), and the step would otherwise make it do that
        in the middle of its first backward, concurrently with the persistent tcgen05 kernels, and again between the
        graph captures of its second.  Ends with a device synchronisation and a barrier: every rank enters its first
        step with all transports up."""
        import torch
        if not self.grad.is_cuda:
            return
        scratch = torch.zeros_like(self.grad)
        for a, b in self.ranges.values():
            if b > a:
                dist.all_reduce(scratch[a:b], op=dist.ReduceOp.SUM, group=self.pg)
        dist.all_reduce(scratch, op=dist.ReduceOp.SUM, group=self.pg)
        torch.cuda.synchronize(self.grad.device)
        dist.barrier(group=self.pg)
        del scratch

    @property
    def active(self):
        return self.world > 1

    def ready(self, tag):
        """The gradient slice `tag` is final on the current stream: start its all-reduce.  torch's NCCL group
        runs the collective on its own stream after the work already enqueued on the current stream, so the
        remaining backward kernels overlap the transfer."""
        if not self.active or tag not in self.ranges or tag in self.done:
            return
        a, b = self.ranges[tag]
        self.done.add(tag)
        if b > a:
            self.pending.append(dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
            self.collectives += 1

    def finish(self):
        """Reduce whatever has not been started (a step whose last backward did not report buckets), wait, and
        return the gradient scale."""
        if not self.active:
            return 1.0
        for tag in BUCKET_ORDER:
            self.ready(tag)
        for w in self.pending:
            w.wait()           # CUDA: makes the current stream wait for the NCCL stream; CPU (gloo): blocks
        self.pending, self.done = [], set()
        return 1.0 / self.world


def reduce_loss_vector(vec, process_group=None):
    """Mean of the per-rank loss vector for logging (the reference gathers pickled dicts to rank 0 and averages:
    aldi/dropin.py:100-118 `_write_metrics`); one tiny all-reduce instead."""
    if process_group is None or dist.get_world_size(process_group) == 1:
        return vec
    out = vec.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=process_group)
    return out / dist.get_world_size(process_group)


def shard_for_rank(items, rank, world):
    """Contiguous per-rank shard of a global batch (D2's TrainingSampler hands every rank its own infinite
    stream; for a fixed synthetic global batch the equivalent is an equal contiguous split)."""
    n = len(items)
    assert n % world == 0, "global batch %d is not divisible by world size %d" % (n, world)
    per = n // world
    return items[rank * per:(rank + 1) * per]


def check_replicas_in_sync(flat, process_group=None, tol=0.0):
    """Debug aid: max |param - rank0 param| over ranks (replicas must stay identical: same reduced gradient, same
    deterministic update)."""
    if process_group is None or dist.get_world_size(process_group) == 1:
        return 0.0
    ref = flat.clone()
    dist.broadcast(ref, src=0, group=process_group)
    d = (flat - ref).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=process_group)
    d = float(d.item())
    assert d <= tol, "data-parallel replicas diverged: max abs diff %g" % d
    return d
