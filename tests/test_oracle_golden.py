"""Pin the oracle's ALDI layer (oracle/aldi_ref.py) against golden vectors produced by the REFERENCE's own
code (tests/golden/make_golden.py imported /root/reference/aldi/*.py).  CPU only."""
import os
import random

import pytest
import torch

import cases
from oracle import aldi_ref, d2_rcnn as d2

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "aldi_golden.pt"), weights_only=False)
SOFT = dict(do_cls_dst=True, do_obj_dst=True, do_rpn_reg_dst=True, do_roih_reg_dst=True)


def close(a, b, rtol=2e-4, atol=1e-6):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a, b)


def check_grads(model, gold, rtol=2e-3):
    got = cases.grad_probe(model, extra=tuple(k for k in gold if k not in cases.GRAD_PROBES))
    for k, g in gold.items():
        if g is None:
            assert got[k] is None, k
            continue
        scale = max(g["abs_sum"], 1e-12)
        assert abs(got[k]["sum"] - g["sum"]) <= rtol * scale, (k, got[k]["sum"], g["sum"])
        assert abs(got[k]["abs_sum"] - g["abs_sum"]) <= rtol * scale, (k, got[k]["abs_sum"], g["abs_sum"])
        close(got[k]["head"], g["head"], rtol=2e-3, atol=1e-6 * scale)


def build(sd, **kw):
    m = aldi_ref.ALDI(num_classes=8, **kw)
    m.load_state_dict(sd, strict=False)
    return m


def test_ema_matches_reference():
    student, teacher_src = cases.ema_modules()
    ema = aldi_ref.EMA(teacher_src, alpha=0.9996, start_iter=1)
    for it, want in enumerate(GOLD["ema"]["states"]):
        cases.ema_perturb_student(student, it)
        ema.update_weights(student, it)
        got = ema.model.state_dict()
        assert set(got) == set(want)
        for k in want:
            assert torch.equal(got[k], want[k]), (it, k)
    # query_embed keys are copied, not averaged (aldi/ema.py:39-41)
    assert torch.equal(ema.model.state_dict()["query_embed.weight"], student.state_dict()["query_embed.weight"])


def test_ema_missing_key_raises():
    student, teacher_src = cases.ema_modules()
    ema = aldi_ref.EMA(teacher_src, alpha=0.9, start_iter=-1)
    del student.conv
    with pytest.raises(Exception):
        ema.update_weights(student, 5)


def test_grad_reverse_matches_reference():
    x = cases.grad_reverse_input()
    y = aldi_ref.grad_reverse(x)
    (y * cases.grad_reverse_weights()).sum().backward()
    assert torch.equal(y.detach(), GOLD["grad_reverse"]["y"])
    assert torch.equal(x.grad, GOLD["grad_reverse"]["grad"])


@pytest.mark.parametrize("name", [n for n, c in cases.CASES.items() if c["kind"] == "distill"])
def test_distiller_matches_reference(name):
    case, gold = cases.CASES[name], GOLD[name]
    sd_s, sd_t = cases.student_teacher_state(case)
    student, teacher = build(sd_s).train(), build(sd_t).train()
    dist = aldi_ref.ALDIDistiller(teacher, student, **SOFT)
    _, uw, us = cases.data(case, d2)
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        losses = dist(uw, us)
        sum(losses.values()).backward()
    assert set(losses) == set(gold["losses"])
    for k in gold["losses"]:
        close(losses[k].detach(), gold["losses"][k])
    for d, g in zip(uw, gold["pseudo"]):
        assert torch.equal(d["instances"].gt_classes, g["classes"])       # bit-exact selection
        close(d["instances"].gt_boxes.tensor, g["boxes"], rtol=1e-5, atol=1e-4)
        close(d["instances"].scores, g["scores"], rtol=1e-5)
    for a, b in zip(uw, us):
        assert a["instances"] is b["instances"]                            # T4
    check_grads(student, gold["grads"])


@pytest.mark.parametrize("name", [n for n, c in cases.CASES.items() if c["kind"] == "train_step"])
def test_train_step_matches_reference(name):
    case, gold = cases.CASES[name], GOLD[name]
    sd_s, sd_t = cases.student_teacher_state(case)
    student, teacher = build(sd_s).train(), build(sd_t).train()
    dist = aldi_ref.ALDIDistiller(teacher, student, **SOFT)
    ls, uw, us = cases.data(case, d2)
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        loss_dict = aldi_ref.run_model_labeled_unlabeled(student, dist, (None, ls, uw, us), case["ims_per_gpu"], False,
                                                         lambda l: l.backward())
    assert set(loss_dict) == set(gold["loss_dict"])
    for k in gold["loss_dict"]:
        close(loss_dict[k], gold["loss_dict"][k])
    check_grads(student, gold["grads"])


def test_align_matches_reference():
    case, gold = cases.CASES["align"], GOLD["align"]
    sd_s, _ = cases.student_teacher_state(case)
    model = build(sd_s, img_da_enabled=True, ins_da_enabled=True).train()
    cases.init_discriminators(model)
    ls, uw, _ = cases.data(case, d2, with_labels_for_unlabeled=True)
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        for tag, batch, labeled in (("labeled", ls, True), ("unlabeled", uw, False)):
            model.zero_grad()
            losses = model(batch, labeled=labeled, do_align=True)
            (losses["loss_da_img"] + losses["loss_da_ins"]).backward()
            close(losses["loss_da_img"].detach(), gold[tag]["loss_da_img"])
            close(losses["loss_da_ins"].detach(), gold[tag]["loss_da_ins"])
            check_grads(model, gold[tag]["grads"])
        losses = model(ls, do_align=False)
        assert float(losses["_da"]) == gold["dummy_da"] == 0.0
