"""Seeded inputs shared by the golden generator (reference side) and the replay tests (oracle side)."""
import torch
from torch import nn

from aldi_b200 import arch, synth_data

CASES = {
    # ALDIDistiller.__call__ on one micro-batch: soft losses + zero-weighted hard losses + pseudo labels
    "distill_2img": {"kind": "distill", "seed": 11, "h": 128, "w": 160, "n_l": 0, "n_u": 2, "teacher_mix": 0.05},
    # teacher == student at init (iter 0 copy): pseudo labels likely fewer / different regime
    "distill_same_teacher": {"kind": "distill", "seed": 12, "h": 96, "w": 128, "n_l": 0, "n_u": 1, "teacher_mix": 0.0},
    # whole run_model_labeled_unlabeled: 2 source + 2 target, micro-batch 2 (clean case of T6)
    "step_2p2_mb2": {"kind": "train_step", "seed": 21, "h": 128, "w": 160, "n_l": 2, "n_u": 2, "ims_per_gpu": 2,
                     "teacher_mix": 0.05},
    # uneven micro-batches (T9): 3 + 3 images with IMS_PER_GPU 2 -> num_grad_accum_steps 3, four micro-batches
    "step_3p3_mb2": {"kind": "train_step", "seed": 22, "h": 96, "w": 96, "n_l": 3, "n_u": 3, "ims_per_gpu": 2,
                     "teacher_mix": 0.05},
    "align": {"kind": "align", "seed": 31, "h": 96, "w": 128, "n_l": 1, "n_u": 1, "teacher_mix": 0.0},
}

GRAD_PROBES = ("backbone.bottom_up.res3.0.conv1.weight", "backbone.bottom_up.res5.2.conv3.weight",
               "backbone.fpn_lateral3.weight", "backbone.fpn_output2.bias",
               "proposal_generator.rpn_head.conv.weight", "proposal_generator.rpn_head.objectness_logits.weight",
               "proposal_generator.rpn_head.anchor_deltas.bias", "roi_heads.box_head.fc1.weight",
               "roi_heads.box_predictor.cls_score.weight", "roi_heads.box_predictor.bbox_pred.weight")


def student_teacher_state(case):
    s = arch.synthetic_state_dict(seed=case["seed"])
    o = arch.synthetic_state_dict(seed=case["seed"] + 1000)
    mix = case["teacher_mix"]
    t = {k: (1 - mix) * s[k] + mix * o[k] for k in s}
    return s, t


def data(case, d2, with_labels_for_unlabeled=False):
    """-> (labeled_strong, unlabeled_weak, unlabeled_strong) as Detectron2-style lists of dicts built on `d2`."""
    ls, uw, us = synth_data.synthetic_batch(case["seed"], case["n_l"], case["n_u"], case["h"], case["w"])

    def conv(d, labeled):
        out = {"image": d["image"].clone(), "height": d["height"], "width": d["width"]}
        if labeled:
            out["instances"] = d2.Instances((d["height"], d["width"]), gt_boxes=d2.Boxes(d["boxes"].clone()),
                                            gt_classes=d["classes"].clone())
        return out

    ls = [conv(d, True) for d in ls]
    uw2 = [conv(d, False) for d in uw]
    us2 = [conv(d, False) for d in us]
    if with_labels_for_unlabeled:
        # the align step runs the full training forward on target images: it needs (any) instances
        gen = torch.Generator().manual_seed(case["seed"] + 5)
        for d in uw2:
            _, b, c = synth_data.synth_image(case["h"], case["w"], gen, 3)
            d["instances"] = d2.Instances((case["h"], case["w"]), gt_boxes=d2.Boxes(b), gt_classes=c)
    return ls, uw2, us2


def grad_probe(model, extra=()):
    sd = dict(model.named_parameters())
    out = {}
    for k in GRAD_PROBES + tuple(extra):
        g = sd[k].grad
        if g is None:
            out[k] = None
            continue
        out[k] = {"sum": g.double().sum().item(), "abs_sum": g.double().abs().sum().item(),
                  "head": g.flatten()[:16].clone()}
    return out


def init_discriminators(model):
    gen = torch.Generator().manual_seed(99)
    for mod in (model.img_align, model.ins_align):
        for p in mod.parameters():
            p.data = torch.randn(p.shape, generator=gen) * 0.02


# ---- EMA toy modules -------------------------------------------------------------------------------
class _Toy(nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.conv = nn.Conv2d(3, 8, 3)
        self.query_embed = nn.Embedding(5, 4)
        self.register_buffer("running_var", torch.rand(8, generator=g) + 0.5)
        for p in self.parameters():
            p.data = torch.randn(p.shape, generator=g)
        self.device = torch.device("cpu")


def ema_modules():
    return _Toy(1), _Toy(2)


def ema_perturb_student(student, it):
    g = torch.Generator().manual_seed(100 + it)
    for p in student.parameters():
        p.data += 0.1 * torch.randn(p.shape, generator=g)
    student.running_var += 0.01


def grad_reverse_input():
    g = torch.Generator().manual_seed(5)
    return torch.randn(2, 3, 4, generator=g).requires_grad_(True)


def grad_reverse_weights():
    g = torch.Generator().manual_seed(6)
    return torch.randn(2, 3, 4, generator=g)
