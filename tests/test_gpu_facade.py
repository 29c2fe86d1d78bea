"""The registry-built facades (aldi_b200/model.py) under the REFERENCE's training loop: `run_model_labeled_unlabeled`
(aldi/trainer.py:28-117, here its pinned restatement oracle/aldi_ref.py) calls `model(batch, labeled=, do_align=)`,
`distiller(teacher_batch, student_batch)` and `do_backward(sum(losses) / num_grad_accum_steps)` exactly as it does on
Detectron2 modules; the loss dict and the accumulated gradient must equal what the engine's own planned step
(`B200TrainStep.run_model`, the benchmarked entry point, unfused here) produces for the same data and seeds."""
import random

import pytest
import torch

import parity_utils as pu
from oracle import aldi_ref, d2_rcnn as d2

pytestmark = pytest.mark.gpu

ALDI_BEST = ("DOMAIN_ADAPT.DISTILL.ROIH_CLS_ENABLED", True, "DOMAIN_ADAPT.DISTILL.OBJ_ENABLED", True,
             "DOMAIN_ADAPT.DISTILL.ROIH_REG_ENABLED", True, "DOMAIN_ADAPT.DISTILL.RPN_REG_ENABLED", True)


def _cfg(opts=(), distiller="ALDIDistiller"):
    from aldi_b200.config import add_aldi_config, get_cfg
    cfg = get_cfg()
    add_aldi_config(cfg)
    # the RPN top-k values of configs/detectron2/Base-RCNN-FPN.yaml:14-20 (the bare defaults are the C4 model's 12000 / 6000)
    cfg.merge_from_list(["MODEL.RPN.PRE_NMS_TOPK_TRAIN", 2000, "MODEL.RPN.PRE_NMS_TOPK_TEST", 1000,
                         "MODEL.RPN.POST_NMS_TOPK_TRAIN", 1000, "MODEL.RPN.POST_NMS_TOPK_TEST", 1000])
    cfg.merge_from_list(["MODEL.ROI_HEADS.NUM_CLASSES", 8, "SOLVER.IMS_PER_GPU", 2, "EMA.ENABLED", True, "EMA.START_ITER", -1,
                         "DOMAIN_ADAPT.DISTILL.DISTILLER_NAME", distiller, "MODEL.DEVICE", "cuda"] + list(opts))
    return cfg


def _engine_step(cfg, sd_s, sd_t, data, seed):
    from aldi_b200.config import step_config_from_cfg
    from aldi_b200.train_step import B200TrainStep
    scfg = step_config_from_cfg(cfg, dtype="fp32")
    scfg.fuse_passes = False
    eng = B200TrainStep(scfg, sd_s, teacher_state_dict=sd_t)
    eng.debug = None
    random.seed(seed)
    losses = dict(eng.run_model(data).items())
    torch.cuda.synchronize()
    return eng, losses


@pytest.mark.parametrize("distiller,opts", [("ALDIDistiller", ALDI_BEST),
                                            ("HardDistiller", ("DOMAIN_ADAPT.DISTILL.HARD_OBJ_ENABLED", True)),
                                            ("Distiller", ())])
def test_reference_loop_over_the_facades_equals_the_engine_step(distiller, opts):
    from aldi_b200 import model as M
    cfg = _cfg(opts, distiller)
    sd_s, sd_t, ls, uw, us = pu.make_inputs(9, 3, 3, 96, 128)           # 3 + 3 images, IMS_PER_GPU 2: the uneven T9 case
    data = (None, ls, uw, us)
    model = M.build_aldi(cfg, state_dict=sd_s, dtype="fp32")
    assert [c.__name__ for c in type(model).__mro__][:4] == ["ALDI", "AlignMixin", "DistillMixin", "GeneralizedRCNN"]
    ema = M.EMA(model, cfg.EMA.ALPHA, cfg.EMA.START_ITER)
    ema.model.load_state_dict(sd_t)
    dist = M.build_distiller(cfg, teacher=ema.model, student=model)
    assert type(dist).__name__ == distiller
    random.seed(77)
    with d2.EventStorage():
        got = aldi_ref.run_model_labeled_unlabeled(model, dist, data, cfg.SOLVER.IMS_PER_GPU, False, lambda l: l.backward())
    torch.cuda.synchronize()
    got = {k: float(v) for k, v in got.items()}
    cfg2 = _cfg(opts, distiller)
    if distiller == "HardDistiller":     # what HardDistiller.engine_flags() switches on: all four standard losses
        cfg2.merge_from_list(["DOMAIN_ADAPT.DISTILL.DISTILLER_NAME", "ALDIDistiller"] + [x for k in (
            "HARD_ROIH_CLS_ENABLED", "HARD_ROIH_REG_ENABLED", "HARD_OBJ_ENABLED", "HARD_RPN_REG_ENABLED")
            for x in ("DOMAIN_ADAPT.DISTILL." + k, True)])
    eng, want = _engine_step(cfg2, sd_s, sd_t, data, 77)
    if distiller == "Distiller":
        assert not any(k.endswith("_distill") for k in got) and not any(k.endswith("_distill") for k in want)
    else:
        assert any(k.endswith("_distill") for k in got)
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-5 * max(abs(want[k]), 1e-3), (k, got[k], want[k])
    assert model.engine.seed_log == eng.seed_log
    g, r = model.engine.grad, eng.grad
    assert float(r.norm()) > 0 and float((g - r).norm() / r.norm()) < 1e-5
    # one optimizer step + EMA through the facades moves both sides of the engine
    before = model.state_dict()["roi_heads.box_head.fc1.weight"].clone()
    model.engine.optimizer_step(lr=0.01)
    ema.update_weights(model, 5)
    assert not torch.equal(model.state_dict()["roi_heads.box_head.fc1.weight"], before)
    inst = ema.model.inference(uw[:1], do_postprocess=False)
    assert len(inst) == 1 and hasattr(inst[0], "pred_boxes")
