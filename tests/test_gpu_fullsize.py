"""Parity at the BASELINE shape: one (source + target) 1024x2048 pair, IMS_PER_GPU 1 (BASELINE.json configs[1] geometry:
523,776 anchors per image, halo-mode 3x3, conv_tn, 256-row wgrad tiles, tap pairing, multi-block radix select, segmented
NMS -- none of which the 128x160 tests reach).

  * test_full_size_step_matches_oracle: the whole ALDI++ step against the CPU oracle (oracle/aldi_ref.py) at 1e-3 on
    every loss and 5e-3 on every parameter gradient, in the CUDA-core fp32 mode AND through the tcgen05 kernels in
    split-bf16 mode (bf16x3: 3 product terms, bf16x6: 6 terms = fp32-level).  Discontinuous selections are compared
    first (the device's own pseudo labels and RPN proposals against the oracle's, as sets) and then pinned to the
    oracle's, so everything downstream sees identical inputs (the same convention as tests/test_gpu_step_parity.py).
  * test_every_tensor_core_launch_of_a_full_size_step_is_within_one_bf16_ulp: the bf16 step exactly as benchmarked
    (fused student pass, every kernel variant picked by these shapes); EVERY aldi_conv_tc / aldi_wgrad_tc launch is
    shadowed by the fp32 CUDA-core kernel on identical operands and must agree to one bf16 ulp of each output element
    (bf16 outputs), 1e-4 (fp32 outputs and weight gradients).
"""
import random

import pytest
import torch

import parity_utils as pu
from oracle import aldi_ref, d2_rcnn as d2

pytestmark = pytest.mark.gpu
H, W = 1024, 2048
RTOL = 1e-3
_CACHE = {}


def _oracle():
    """One CPU oracle step (~10-20 s), shared by the parametrised device runs."""
    if "ora" in _CACHE:
        return _CACHE["ora"]
    sd_s, sd_t, ls, uw, us = pu.make_inputs(91, 1, 1, H, W)
    pu.install_device_sampler(pu.predict_seed_log(1234, 1, 1))
    student, teacher = pu.oracle_models(sd_s, sd_t)
    dist = aldi_ref.ALDIDistiller(teacher, student, **pu.SOFT)
    prop_log = []
    student.proposal_generator.register_forward_hook(
        lambda m, i, o: prop_log.append([p.proposal_boxes.tensor.detach().clone() for p in o[0]]))
    uw_o, us_o = pu.to_d2(uw, False), pu.to_d2(us, False)
    with d2.EventStorage():
        losses = aldi_ref.run_model_labeled_unlabeled(student, dist, (None, pu.to_d2(ls, True), uw_o, us_o), 1, False,
                                                      lambda l: l.backward())
    d2.set_sample_chooser(None)
    assert len(prop_log) == 2          # student forward of the source pass, then of the distillation pass
    _CACHE["ora"] = dict(sd_s=sd_s, sd_t=sd_t, data=(None, ls, uw, us), losses={k: float(v) for k, v in losses.items()},
                         student=student, dist=dist, uw_o=uw_o, props={0: prop_log[0], 101: prop_log[1]})
    return _CACHE["ora"]


def _set_agreement(a, b, tol=0.05):
    """fraction of boxes of `a` that have a partner in `b` within `tol` pixels (max-abs over the 4 coordinates)."""
    if a.shape[0] == 0 or b.shape[0] == 0:
        return 1.0 if a.shape[0] == b.shape[0] else 0.0
    d = (a[:, None, :] - b[None, :, :]).abs().amax(dim=2)
    return float((d.amin(dim=1) <= tol).float().mean())


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("dtype", ["fp32", "bf16x3", "bf16x6"])
def test_full_size_step_matches_oracle(dtype):
    from aldi_b200.train_step import B200TrainStep, StepConfig
    o = _oracle()
    step = B200TrainStep(StepConfig(dtype=dtype, ims_per_gpu=1, ema_start_iter=-1), o["sd_s"], teacher_state_dict=o["sd_t"])
    step.pseudo_override = [pu.pseudo_to_device([o["uw_o"][0]["instances"]], "cuda")]
    step.proposal_override = o["props"]
    random.seed(1234)
    dev = dict(step.run_model(o["data"]).items())
    torch.cuda.synchronize()
    assert step.seed_log == pu.predict_seed_log(1234, 1, 1)
    # --- the device's own selections against the oracle's (sets; see the module docstring)
    pseudo, inst = step.pseudo_log[-1], o["uw_o"][0]["instances"]
    k = int(pseudo.counts[0])
    assert k == len(inst), ("pseudo-label count", k, len(inst))
    if k:
        assert torch.equal(pseudo.classes[0, :k].cpu().long(), inst.gt_classes), "pseudo-label classes"
        assert torch.allclose(pseudo.boxes[0, :k].cpu(), inst.gt_boxes.tensor, rtol=1e-4, atol=5e-2)
        assert torch.allclose(pseudo.scores[0, :k].cpu(), inst.scores, rtol=1e-3, atol=1e-4)
    agree = {}
    for pid, want in o["props"].items():
        boxes, count = step.proposal_log[pid]
        got = boxes[0, :int(count[0])].cpu()
        agree[pid] = (_set_agreement(got, want[0]), _set_agreement(want[0], got), got.shape[0], want[0].shape[0])
        assert abs(got.shape[0] - want[0].shape[0]) <= 10 and min(agree[pid][:2]) >= 0.98, ("RPN proposals", pid, agree[pid])
    # --- sampled anchors of the distillation loss (inputs pinned -> exact), then losses and every gradient
    got_l, want_l = step.debug["labels"].cpu().to(torch.int8), o["dist"].io["distill_labels"].to(torch.int8)
    assert int((got_l != want_l).sum()) == 0, "distillation anchor labels"
    for kk, v in o["losses"].items():
        assert abs(dev[kk] - v) <= RTOL * max(abs(v), 1e-3), (dtype, kk, dev[kk], v)
    g = step.grad.cpu()
    worst = ("", 0.0)
    for key, (off, n, ref) in pu.oracle_grads_internal(step.layout, o["student"]).items():
        if ref is None:
            assert float(g[off:off + n].abs().max()) == 0.0, key
            continue
        e = pu.rel_err(g[off:off + n], ref)
        worst = max(worst, (key, e), key=lambda t: t[1])
        assert e < 5e-3, (dtype, key, e)
    print("full-size %s: losses %s\n  proposal agreement %s, worst gradient rel err %s" % (dtype, dev, agree, worst))


@pytest.mark.timeout(1200)
def test_every_tensor_core_launch_of_a_full_size_step_is_within_one_bf16_ulp():
    from aldi_b200 import ops
    from aldi_b200.train_step import B200TrainStep, StepConfig
    sd_s, sd_t, ls, uw, us = pu.make_inputs(91, 1, 1, H, W)
    step = B200TrainStep(StepConfig(dtype="bf16", ims_per_gpu=1, ema_start_iter=-1), sd_s, teacher_state_dict=sd_t)
    seen, worst = {}, {"conv": ("", 0.0), "wgrad": ("", 0.0)}

    def shadow(kind, run, **c):
        kw = dict(c["kw"])
        key = kind + " " + c["key"]
        if kind == "conv":
            out, cs = c["out"], kw["cout_store"]
            ref = out.float() if kw["accumulate"] else torch.zeros(out.shape, device=out.device)
            run()
            for name in ("residual", "mask"):
                if kw[name] is not None:
                    kw[name] = kw[name].float()
            ops.conv(c["x"].float(), c["wp"].float(), ref, **kw)          # fp32 operands -> aldi_conv_f32
            got, want = out[..., :cs].float(), ref[..., :cs]
            rms = float(want.pow(2).mean().sqrt())
            rel = 2.0 ** -8 if out.dtype == torch.bfloat16 else 1e-4
            # + the fp32 summation-order noise of the two kernels (K up to 12544 products per output): 1e-4 of the rms
            ratio = float(((got - want).abs() / (rel * want.abs() + 1e-4 * rms + 1e-30)).max())
        else:
            dw = c["dw"]
            before = dw.clone()
            run()
            ref = torch.zeros_like(dw)
            ops.wgrad(c["x"].float(), c["dy"].float(), ref, **kw)          # -> aldi_wgrad_f32
            delta = dw - before
            bound = 1e-4 * float(ref.abs().max()) + 1e-6 * float(before.abs().max()) + 1e-30
            ratio = float((delta - ref).abs().max()) / bound
        seen[key] = max(seen.get(key, 0.0), ratio)
        if ratio > worst[kind][1]:
            worst[kind] = (key, ratio)

    ops.set_shadow(shadow)
    try:
        random.seed(1234)
        losses = dict(step.run_model((None, ls, uw, us)).items())
        torch.cuda.synchronize()
    finally:
        ops.set_shadow(None)
    kinds = {k.split(" ")[0] for k in seen}
    assert kinds == {"conv", "wgrad"} and len(seen) >= 80, (len(seen), kinds)
    bad = {k: v for k, v in seen.items() if not v <= 1.0}
    print("shadowed %d distinct tensor-core layer shapes; worst (error / bound): %s; losses %s" % (len(seen), worst, losses))
    assert not bad, bad
