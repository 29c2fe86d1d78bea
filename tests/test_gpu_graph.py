"""CUDA-graph replay of the micro-batches must reproduce the eager schedule: same kernels, same inputs (seed and
salts are device-resident), so the losses of a step agree exactly as long as the weights do."""
import random

import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu


def _steps(cuda_graph, n_steps, dtype):
    from aldi_b200.train_step import B200TrainStep, StepConfig
    sd_s, sd_t, ls, uw, us = pu.make_inputs(5, 2, 2, 128, 160)
    step = B200TrainStep(StepConfig(dtype=dtype, ims_per_gpu=2, ema_start_iter=-1, cuda_graph=cuda_graph, base_lr=1e-4),
                         sd_s, teacher_state_dict=sd_t)
    step.debug = None
    random.seed(7)
    out = []
    for _ in range(n_steps):
        # lr = 0: the weights stay put (the EMA teacher still moves, deterministically), so eager and replayed steps
        # see bit-identical operands and only the fp32 atomics inside the loss reductions may reorder
        losses = step.step((None, ls, uw, us), lr=0.0)
        out.append(dict(losses.items()))
    torch.cuda.synchronize()
    return step, out


@pytest.mark.parametrize("dtype,split", [("fp32", False), ("bf16", False), ("bf16", True)])
def test_graph_replay_matches_eager(dtype, split, monkeypatch):
    """split: the last backward captured as a chain of graphs cut at the gradient-bucket boundaries (what a
    data-parallel run does so that each bucket's all-reduce can start between two replays)."""
    if split:
        monkeypatch.setenv("ALDI_GRAPH_SPLIT", "1")
    eager_step, eager = _steps(False, 4, dtype)
    graph_step, graph = _steps(True, 4, dtype)
    assert eager_step.graph_replays == 0
    # (source, distill) x steps 2..4, step 1 runs eagerly; split: the distill backward is 6 segments
    assert graph_step.graph_replays == (1 + 6 if split else 2) * 3
    for i, (a, b) in enumerate(zip(eager, graph)):
        assert set(a) == set(b)
        for k in a:
            tol = 1e-5
            assert abs(a[k] - b[k]) <= tol * max(abs(a[k]), 1e-3), (i, k, a[k], b[k])
    # sampled sets are seed-driven and the seed lives in device memory: a replay must draw NEW samples each step
    assert graph[1] != graph[2]


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """aldi_b200/checkpoint.py through the real step: save after one optimizer step, load into a fresh step built from
    other weights, and the flat student / teacher / momentum buffers are identical; a fresh (non-resume) start from
    the same file takes the EMA weights as the model (aldi/checkpoint.py:19-31)."""
    from aldi_b200 import arch
    from aldi_b200.checkpoint import DetectionCheckpointerWithEMA
    from aldi_b200.train_step import B200TrainStep, StepConfig
    sd_s, sd_t, ls, uw, us = pu.make_inputs(5, 2, 2, 96, 128)
    cfg = StepConfig(dtype="bf16", ims_per_gpu=2, ema_start_iter=-1, base_lr=1e-3)
    a = B200TrainStep(cfg, sd_s, teacher_state_dict=sd_t)
    a.debug = None
    random.seed(3)
    a.step((None, ls, uw, us))
    torch.cuda.synchronize()
    path = DetectionCheckpointerWithEMA(a, str(tmp_path)).save("model_0000000")
    saved_teacher = a.teacher.flat.clone()
    other = arch.synthetic_state_dict(seed=77)
    b = B200TrainStep(cfg, other)
    b.debug = None
    DetectionCheckpointerWithEMA(b, str(tmp_path)).resume_or_load("", resume=True)
    assert torch.equal(a.student.flat, b.student.flat)
    assert torch.equal(a.teacher.flat, b.teacher.flat)
    assert torch.equal(a.momentum_buf, b.momentum_buf) and b.iter == a.iter == 1
    # the GEMM operands were re-derived from the loaded master weights: the next step agrees
    random.seed(4)
    la = dict(a.step((None, ls, uw, us), lr=0.0).items())
    random.seed(4)
    lb = dict(b.step((None, ls, uw, us), lr=0.0).items())
    for k in la:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(abs(la[k]), 1e-3), (k, la[k], lb[k])
    c = B200TrainStep(cfg, other)
    DetectionCheckpointerWithEMA(c, str(tmp_path / "fresh")).resume_or_load(path, resume=False)
    nt = c.layout.numel
    assert torch.equal(c.student.flat[:nt], saved_teacher[:nt])


def test_convnext_step_replays_as_graphs_with_fresh_droppath_masks():
    """BASELINE configs[4] inside the graph machinery: the DropPath factors are drawn on the host per forward (as
    torch.bernoulli_ does in the reference, aldi/backbone.py:176-181) but live in device blocks refreshed from pinned
    memory by a captured copy, so the ConvNeXt bodies replay as CUDA graphs and every replay sees NEW masks."""
    from aldi_b200 import arch, synth_data
    from aldi_b200.convnext import synthetic_state_dict as convnext_init
    from aldi_b200.train_step import B200TrainStep, StepConfig
    depths, dims = (1, 1, 2, 1), (32, 64, 96, 128)
    sd = arch.synthetic_state_dict(3, bottom_up_channels=dims)
    sd.update({"backbone.bottom_up." + k: v for k, v in convnext_init(depths, dims, 3, 1.0).items()})
    ls, uw, us = synth_data.synthetic_batch(3, 2, 2, 96, 128)
    cfg = StepConfig(dtype="bf16", ims_per_gpu=2, ema_start_iter=-1, backbone="convnext", convnext_depths=depths,
                     convnext_dims=dims, convnext_drop_path=0.5, optimizer="ADAMW", base_lr=1e-5, cuda_graph=True)
    step = B200TrainStep(cfg, sd)
    step.debug = None
    random.seed(5)
    seen, losses = [], []
    for _ in range(5):
        losses.append(dict(step.step((None, ls, uw, us)).items()))
        torch.cuda.synchronize()
        seen.append(torch.cat([dev.flatten() for slots in step._mask_slots.values() for _, dev, _ in slots]).cpu())
    assert step.graph_replays > 0 and len(step._mask_slots) >= 2          # source body + distillation body
    assert all(v == v and abs(v) < 1e6 for l in losses for v in l.values()), losses[-1]
    # replays (steps 3..5) drew new masks each time; values are 0 or 1 / keep_prob
    assert not torch.equal(seen[2], seen[3]) or not torch.equal(seen[3], seen[4])
    vals = set(torch.cat(seen).unique().tolist())
    assert len(vals) > 1 and 0.0 in vals and all(v == 0.0 or v > 1.0 for v in vals), vals
