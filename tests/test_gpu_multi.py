"""Data-parallel train step on real GPUs (aldi/dropin.py:53,84-85): one process per GPU over NCCL.

Skipped on boxes with fewer than two devices.  What is held:
  * the bucketed all-reduce issued during the last backward leaves, on every rank, the SUM of the gradients the
    ranks compute on their own (each rank also runs the same micro-batches through a step without a process group);
  * after several optimizer steps with CUDA graphs on (segmented capture of the last backward, all-reduces between
    the replays) the student and teacher replicas are bit-identical across ranks;
  * five collectives per step, graph replays actually happened, losses finite.
"""
import os
import random
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist

    import parity_utils as pu
    from aldi_b200 import data_parallel as dp
    from aldi_b200 import synth_data
    from aldi_b200.train_step import B200TrainStep, StepConfig
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        h, w = 128, 160
        sd_s, sd_t, _, _, _ = pu.make_inputs(5, 2, 2, h, w)
        ls, uw, us = synth_data.synthetic_batch(50 + rank, 2, 2, h, w)     # every rank its own images
        data = (None, ls, uw, us)

        def make(pg, graph):
            st = B200TrainStep(StepConfig(dtype="bf16", ims_per_gpu=2, ema_start_iter=-1, cuda_graph=graph, base_lr=1e-3),
                               sd_s, teacher_state_dict=sd_t, device=device, process_group=pg)
            st.debug = None
            return st

        # ---- reduced gradient == sum over ranks of the single-rank gradients
        solo, par = make(None, False), make(dist.group.WORLD, False)
        for st in (solo, par):
            random.seed(11)
            st.ema_update(0)
            st.run_model(data)
        scale = par.allreduce_grads()
        assert scale == 1.0 / world and par.reducer.collectives == 5
        ref = solo.grad.clone()
        dist.all_reduce(ref)
        err = float((par.grad - ref).norm() / ref.norm())
        # two runs of the same bf16 step differ by the order of the fp32 atomics (weight-gradient split-K, RoIAlign
        # backward) and by the bf16 roundings downstream of those sums: ~1e-4 of the gradient norm measured
        assert float(ref.norm()) > 0 and err < 2e-3, ("reduced gradient vs sum of single-rank gradients", err)
        # a rank's own gradient is NOT the reduced one (the ranks really saw different data)
        assert float((solo.grad * world - ref).norm() / ref.norm()) > 1e-2

        # ---- replicas stay bit-identical over optimizer steps with graphs on
        st = make(dist.group.WORLD, True)
        random.seed(100 + rank)            # per-rank sampling seeds, as independent trainer processes have
        last = None
        for _ in range(4):
            last = dict(st.step(data).items())
        torch.cuda.synchronize()
        assert st.graph_replays > 0 and st.reducer.collectives == 5 * 4
        assert all(v == v and abs(v) < 1e6 for v in last.values()), last
        assert dp.check_replicas_in_sync(st.student.flat, dist.group.WORLD) == 0.0
        assert dp.check_replicas_in_sync(st.teacher.flat, dist.group.WORLD) == 0.0
        assert dp.check_replicas_in_sync(st.momentum_buf, dist.group.WORLD) == 0.0
        # ... and they moved: the step was not a no-op
        assert float((st.student.flat[:st.nt] - solo.student.flat[:st.nt]).abs().max()) > 0
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 8])
def test_data_parallel_step_on_gpus(world):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d CUDA devices" % world)
    port = _free_port()
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {r: "ok" for r in range(world)}
