"""Shared helpers for the GPU parity tests: run the CPU oracle and the B200 step on identical inputs."""
import random

import torch

from aldi_b200 import arch, sampling, synth_data
from oracle import aldi_ref, d2_rcnn as d2

SITE = {"rpn": sampling.SITE_RPN, "roi": sampling.SITE_ROI, "rpn_distill": sampling.SITE_RPN_DISTILL}
SOFT = dict(do_cls_dst=True, do_obj_dst=True, do_rpn_reg_dst=True, do_roih_reg_dst=True)


def install_device_sampler(seed_log):
    """Make the oracle draw the samples the device's hash sampler draws (see aldi_b200/sampling.py)."""

    def chooser(candidates, take, tag, ctx):
        salt = sampling.make_salt(ctx["pass"], SITE[ctx["site"]], ctx["image"])
        pos = sampling.choose(seed_log[ctx["pass"]], salt, tag == "neg", candidates.cpu().numpy(), take)
        return torch.from_numpy(pos)

    d2.set_sample_chooser(chooser)


ALIGN = {"img": (256, (256,)), "ins": (1024, (1024,))}


def make_inputs(seed, n_l, n_u, h, w, teacher_mix=0.05, align=None):
    s = arch.synthetic_state_dict(seed=seed, align=align)
    o = arch.synthetic_state_dict(seed=seed + 1000, align=align)
    t = {k: (1 - teacher_mix) * s[k] + teacher_mix * o[k] for k in s}
    ls, uw, us = synth_data.synthetic_batch(seed, n_l, n_u, h, w)
    return s, t, ls, uw, us


def to_d2(batch, labeled, empty_instances=False):
    """empty_instances: what the reference's UnlabeledDatasetMapper attaches to target images
    (aldi/dataloader.py:21-30) — needed whenever the student runs in training mode on them before pseudo-labelling."""
    out = []
    for d in batch:
        e = {"image": d["image"].clone(), "height": d["height"], "width": d["width"]}
        if labeled:
            e["instances"] = d2.Instances((d["height"], d["width"]), gt_boxes=d2.Boxes(d["boxes"].clone()),
                                          gt_classes=d["classes"].clone())
        elif empty_instances:
            e["instances"] = d2.Instances((d["height"], d["width"]), gt_boxes=d2.Boxes(torch.zeros(0, 4)),
                                          gt_classes=torch.zeros(0, dtype=torch.int64))
        out.append(e)
    return out


def oracle_models(sd_s, sd_t, **kw):
    student, teacher = aldi_ref.ALDI(num_classes=8, **kw), aldi_ref.ALDI(num_classes=8, **kw)
    student.load_state_dict(sd_s)
    teacher.load_state_dict(sd_t)
    return student.train(), teacher.train()


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rpn_out_to_d2(rpn_out, lv, n):
    """device (N, total_locs, 16) -> detectron2 head outputs ([N,A,H,W] logits, [N,4A,H,W] deltas) per level."""
    logits, deltas = [], []
    r = rpn_out.cpu()
    for i in range(lv.num_levels):
        h, w, off = lv.h[i], lv.w[i], lv.loc_off[i]
        blk = r[:, off:off + h * w, :].reshape(n, h, w, -1)
        logits.append(blk[..., :3].permute(0, 3, 1, 2).contiguous())
        deltas.append(blk[..., 3:15].permute(0, 3, 1, 2).contiguous())
    return logits, deltas


def oracle_grads_internal(layout, model):
    """oracle parameter gradients -> the flat-buffer element order of the device gradient."""
    named = dict(model.named_parameters())
    out = {}
    for (layer, field), (off, n, key, shape) in layout.entries.items():
        if off >= layout.num_trainable or field not in ("weight", "bias"):
            continue
        g = named[key].grad
        out[key] = (off, n, None if g is None else layout.to_internal(layer, field, g))
    return out


def predict_seed_log(py_seed, n_source_mb, n_distill_mb, n_align_mb=0):
    """The sampling seeds B200TrainStep.run_model will draw after random.seed(py_seed)."""
    rng = random.Random(py_seed)
    log = {}
    s0 = rng.randint(0, 2 ** 32 - 1)
    p = 0
    for _ in range(n_source_mb + n_align_mb):
        log[p] = s0
        p += 1
    for _ in range(n_distill_mb):
        log[100 + p] = rng.randint(0, 2 ** 32 - 1)
        p += 1
    return log


def pseudo_to_device(insts, device):
    """oracle pseudo-label Instances of one micro-batch -> aldi_b200.train_step.GroundTruth."""
    from aldi_b200.train_step import GroundTruth
    n, gmax = len(insts), 128
    b = torch.zeros(n, gmax, 4)
    c = torch.zeros(n, gmax, dtype=torch.int32)
    sc = torch.zeros(n, gmax)
    cnt = torch.zeros(n, dtype=torch.int32)
    for i, inst in enumerate(insts):
        k = len(inst)
        b[i, :k] = inst.gt_boxes.tensor
        c[i, :k] = inst.gt_classes.int()
        sc[i, :k] = inst.scores
        cnt[i] = k
    return GroundTruth(b.to(device), c.to(device), cnt.to(device), gmax, sc.to(device))


def log_student_proposals(student):
    """Forward hook on the oracle student's RPN: one entry per forward = [proposal boxes per image]."""
    log = []
    student.proposal_generator.register_forward_hook(
        lambda m, i, o: log.append([p.proposal_boxes.tensor.detach().clone() for p in o[0]]))
    return log


def proposal_override(log, n_source_mb, n_distill_mb, n_align_mb=0):
    """-> B200TrainStep.proposal_override: the oracle's proposals keyed by the step's pass ids (student forwards run in
    plan order: source micro-batches, alignment passes, then the distillation micro-batches as 100 + index)."""
    ids = list(range(n_source_mb + n_align_mb)) + [100 + n_source_mb + n_align_mb + j for j in range(n_distill_mb)]
    assert len(log) == len(ids), (len(log), ids)
    return dict(zip(ids, log))
