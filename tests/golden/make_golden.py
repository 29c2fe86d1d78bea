"""Generate golden vectors from the REFERENCE's own Python code (run in the authoring container only).

The real modules /root/reference/aldi/{helpers,ema,pseudolabeler,distill,align,model,trainer}.py are
imported with oracle/d2shim.py standing in for the uninstallable Detectron2 (whose Faster R-CNN is restated
in oracle/d2_rcnn.py).  Outputs are committed as tests/golden/aldi_golden.pt and replayed against the
oracle's ALDI restatement (oracle/aldi_ref.py) by tests/test_oracle_golden.py; inputs are regenerated from
seeds by tests/golden/cases.py so the fixture stays small.

    python tests/golden/make_golden.py
"""
import os
import random
import sys
import types

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))

from oracle import d2shim  # noqa: E402

d2shim.install()

import aldi.align  # noqa: E402,F401
import aldi.distill  # noqa: E402
import aldi.ema  # noqa: E402
import aldi.helpers  # noqa: E402
import aldi.model  # noqa: E402
import aldi.trainer  # noqa: E402
from aldi.config import add_aldi_config  # noqa: E402

import cases  # noqa: E402
from oracle import d2_rcnn as d2  # noqa: E402


def build_reference_model(sd, align=False):
    cfg = d2shim.get_cfg()
    add_aldi_config(cfg)
    if align:
        cfg.DOMAIN_ADAPT.ALIGN.IMG_DA_ENABLED = True
        cfg.DOMAIN_ADAPT.ALIGN.INS_DA_ENABLED = True
    model = aldi.model.build_aldi(cfg)
    model.load_state_dict(sd, strict=not align)
    if align:
        cases.init_discriminators(model)
    return model, cfg


def reference_distiller(cfg, teacher, student):
    for k in ("ROIH_CLS_ENABLED", "ROIH_REG_ENABLED", "OBJ_ENABLED", "RPN_REG_ENABLED"):
        cfg.DOMAIN_ADAPT.DISTILL[k] = True
    return aldi.distill.ALDIDistiller.from_config(cfg, teacher, student)


def golden_ema():
    out = {}
    student, teacher_src = cases.ema_modules()
    ema = aldi.ema.EMA(teacher_src, alpha=0.9996, start_iter=1)
    states = []
    for it in range(4):
        cases.ema_perturb_student(student, it)
        ema.update_weights(student, it)
        states.append({k: v.clone() for k, v in ema.model.state_dict().items()})
    out["states"] = states
    return out


def golden_grad_reverse():
    x = cases.grad_reverse_input()
    y = aldi.helpers.grad_reverse(x)
    (y * cases.grad_reverse_weights()).sum().backward()
    return {"y": y.detach().clone(), "grad": x.grad.clone()}


def golden_distill(case):
    sd_s, sd_t = cases.student_teacher_state(case)
    student, cfg = build_reference_model(sd_s)
    teacher, _ = build_reference_model(sd_t)
    student.train(); teacher.train()
    distiller = reference_distiller(cfg, teacher, student)
    _, uw, us = cases.data(case, d2)
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        losses = distiller(uw, us)
        total = sum(losses.values())
        total.backward()
    return {
        "losses": {k: v.detach().clone() for k, v in losses.items()},
        "pseudo": [{"boxes": d["instances"].gt_boxes.tensor.clone(), "classes": d["instances"].gt_classes.clone(),
                    "scores": d["instances"].scores.clone()} for d in uw],
        "grads": cases.grad_probe(student),
    }


def golden_train_step(case):
    sd_s, sd_t = cases.student_teacher_state(case)
    student, cfg = build_reference_model(sd_s)
    teacher, _ = build_reference_model(sd_t)
    student.train(); teacher.train()
    distiller = reference_distiller(cfg, teacher, student)
    ls, uw, us = cases.data(case, d2)
    trainer = types.SimpleNamespace(model=student, backward_at_end=False, model_batch_size=case["ims_per_gpu"],
                                    distiller=distiller,
                                    do_backward=lambda losses, override=False: losses.backward())
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        loss_dict = aldi.trainer.run_model_labeled_unlabeled(trainer, None, ls, uw, us)
    return {"loss_dict": {k: v.detach().clone() for k, v in loss_dict.items()}, "grads": cases.grad_probe(student)}


def golden_align(case):
    sd_s, _ = cases.student_teacher_state(case)
    model, cfg = build_reference_model(sd_s, align=True)
    model.train()
    ls, uw, _ = cases.data(case, d2, with_labels_for_unlabeled=True)
    out = {}
    random.seed(case["seed"]); torch.manual_seed(case["seed"])
    with d2.EventStorage():
        for tag, batch, labeled in (("labeled", ls, True), ("unlabeled", uw, False)):
            model.zero_grad()
            losses = model(batch, labeled=labeled, do_align=True)
            (losses["loss_da_img"] + losses["loss_da_ins"]).backward()
            out[tag] = {"loss_da_img": losses["loss_da_img"].detach().clone(),
                        "loss_da_ins": losses["loss_da_ins"].detach().clone(),
                        "grads": cases.grad_probe(model, extra=("img_align.model.0.weight", "ins_align.model.1.weight"))}
        losses = model(ls, do_align=False)
        out["dummy_da"] = float(losses["_da"])
    return out


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g = {"ema": golden_ema(), "grad_reverse": golden_grad_reverse()}
    for name, case in cases.CASES.items():
        kind = case["kind"]
        print("generating", name, flush=True)
        if kind == "distill":
            g[name] = golden_distill(case)
        elif kind == "train_step":
            g[name] = golden_train_step(case)
        elif kind == "align":
            g[name] = golden_align(case)
    path = os.path.join(os.path.dirname(__file__), "aldi_golden.pt")
    torch.save(g, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for name in g:
        if "losses" in g[name]:
            print(name, {k: float(v) for k, v in g[name]["losses"].items()}, [len(p["scores"]) for p in g[name]["pseudo"]])
        if "loss_dict" in g[name]:
            print(name, {k: float(v) for k, v in g[name]["loss_dict"].items()})


if __name__ == "__main__":
    main()
