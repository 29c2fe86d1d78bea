"""ORACLE (test infrastructure, not product code): CPU restatement of the ALDI layer of the hot path —
the model mixins, the distiller, the pseudo-labeler, the EMA teacher and the train-step orchestration —
on top of the Detectron2 restatement in oracle/d2_rcnn.py.

Each function cites the reference lines it follows (paths relative to the reference root).  Unlike
oracle/d2_rcnn.py this layer IS pinned: tests/golden/make_golden.py imports the real reference modules
(aldi/distill.py, aldi/ema.py, aldi/pseudolabeler.py, aldi/trainer.py, aldi/align.py, aldi/helpers.py) in the
authoring container, runs them over the same Detectron2 restatement, and commits the outputs as golden
fixtures that tests/test_oracle_golden.py replays against this file.
"""
import copy
import random
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

from . import d2_rcnn as d2


# ---------------------------------------------------------------------------------------------------
# aldi/helpers.py:51-63  gradient reversal
# ---------------------------------------------------------------------------------------------------
class _GradScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        ctx.weight = weight
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.weight * g.clone(), None


def grad_reverse(x):
    return _GradScale.apply(x, -1.0)


# ---------------------------------------------------------------------------------------------------
# aldi/align.py:103-136  discriminators
# ---------------------------------------------------------------------------------------------------
class ConvDiscriminator(nn.Module):
    def __init__(self, input_dim, hidden_dims=(), kernel_size=3):
        super().__init__()
        mods, prev = [], input_dim
        for dim in hidden_dims:
            mods += [nn.Conv2d(prev, dim, kernel_size), nn.ReLU()]
            prev = dim
        mods += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(prev, 1)]
        self.model = nn.Sequential(*mods)

    def forward(self, x):
        return self.model(x)


class FCDiscriminator(nn.Module):
    def __init__(self, input_dim, hidden_dims=()):
        super().__init__()
        mods, prev = [nn.Flatten()], input_dim
        for dim in hidden_dims:
            mods += [nn.Linear(prev, dim), nn.ReLU()]
            prev = dim
        mods.append(nn.Linear(prev, 1))
        self.model = nn.Sequential(*mods)

    def forward(self, x):
        return self.model(x)


# ---------------------------------------------------------------------------------------------------
# aldi/model.py:12-34 + aldi/align.py:17-101: class ALDI(AlignMixin, DistillMixin, GeneralizedRCNN)
# ---------------------------------------------------------------------------------------------------
class ALDI(d2.GeneralizedRCNN):
    def __init__(self, *, img_da_enabled=False, img_da_layer="p2", img_da_weight=0.01, img_da_input_dim=256,
                 img_da_hidden_dims=(256,), ins_da_enabled=False, ins_da_weight=0.01, ins_da_input_dim=1024,
                 ins_da_hidden_dims=(1024,), **kwargs):
        super().__init__(**kwargs)
        self.img_da_layer, self.img_da_weight, self.ins_da_weight = img_da_layer, img_da_weight, ins_da_weight
        self.img_align = ConvDiscriminator(img_da_input_dim, img_da_hidden_dims) if img_da_enabled else None
        self.ins_align = FCDiscriminator(ins_da_input_dim, ins_da_hidden_dims) if ins_da_enabled else None
        # aldi/align.py:44-52: outputs of sub-modules are captured by forward hooks
        self.taps = {}
        self.backbone.register_forward_hook(lambda m, i, o: self.taps.__setitem__("backbone", o))
        self.roi_heads.box_head.register_forward_hook(lambda m, i, o: self.taps.__setitem__("box_head", o))

    def forward(self, batched_inputs, labeled=True, do_align=False):
        output = super().forward(batched_inputs)
        if self.training:
            if do_align:  # aldi/align.py:74-90
                domain_label = 1 if labeled else 0
                if self.img_align:
                    feats = grad_reverse(self.taps["backbone"][self.img_da_layer])
                    preds = self.img_align(feats)
                    loss = F.binary_cross_entropy_with_logits(preds, torch.full_like(preds, float(domain_label)))
                    output["loss_da_img"] = self.img_da_weight * loss
                if self.ins_align:
                    feats = grad_reverse(self.taps["box_head"])
                    preds = self.ins_align(feats)
                    loss = F.binary_cross_entropy_with_logits(preds, torch.full_like(preds, float(domain_label)))
                    output["loss_da_ins"] = self.ins_da_weight * loss
            elif self.img_align or self.ins_align:  # aldi/align.py:91-100
                fake = 0
                for aligner in [self.img_align, self.ins_align]:
                    if aligner is not None:
                        fake += sum([p.sum() for p in aligner.parameters()]) * 0
                output["_da"] = fake
        return output


# ---------------------------------------------------------------------------------------------------
# aldi/ema.py:8-60
# ---------------------------------------------------------------------------------------------------
class EMA(nn.Module):
    def __init__(self, model, alpha, start_iter=0):
        super().__init__()
        self.model = copy.deepcopy(model)
        self.alpha = alpha
        self.start_iter = start_iter
        self.exclude_keys = ["query_embed"]

    def update_weights(self, model, iter):
        student = model.state_dict()
        if iter <= self.start_iter:  # aldi/ema.py:54-55
            self.model.load_state_dict(student)
            return
        new = OrderedDict()  # aldi/ema.py:32-50
        for key, value in self.model.state_dict().items():
            if key not in student:
                raise Exception("{} is not found in student model".format(key))
            if any(k in key for k in self.exclude_keys):
                new[key] = student[key] * 1
            else:
                new[key] = student[key] * (1 - self.alpha) + value * self.alpha
        self.model.load_state_dict(new)


# ---------------------------------------------------------------------------------------------------
# aldi/pseudolabeler.py:15-73
# ---------------------------------------------------------------------------------------------------
def process_bbox(inst, thres):
    valid = inst.scores > thres
    new = d2.Instances(inst.image_size)
    new.gt_boxes = d2.Boxes(inst.pred_boxes.tensor[valid, :]).to("cpu")
    new.gt_classes = inst.pred_classes[valid].to("cpu")
    new.scores = inst.scores[valid].to("cpu")
    return new


def pseudo_label_inplace(model, unlabeled_weak, unlabeled_strong, threshold):
    with torch.no_grad():
        was_training = model.training
        model.eval()
        preds = model.inference(unlabeled_weak, do_postprocess=False)
        if was_training:
            model.train()
        labels = [process_bbox(p, threshold) for p in preds]
        for datum, lab in zip(unlabeled_weak, labels):
            datum["instances"] = lab
        if unlabeled_strong is not None:
            for datum, lab in zip(unlabeled_strong, labels):  # T4: the SAME Instances object
                datum["instances"] = lab


# ---------------------------------------------------------------------------------------------------
# aldi/distill.py:87-278  ALDIDistiller
# ---------------------------------------------------------------------------------------------------
class ALDIDistiller:
    def __init__(self, teacher, student, do_hard_cls=False, do_hard_obj=False, do_hard_rpn_reg=False,
                 do_hard_roi_reg=False, do_cls_dst=False, do_obj_dst=False, do_rpn_reg_dst=False,
                 do_roih_reg_dst=False, cls_temperature=1.0, obj_temperature=1.0, cls_loss_type="CE",
                 pseudo_label_threshold=0.8):
        self.teacher, self.student = teacher, student
        self.do_hard_cls, self.do_hard_obj = do_hard_cls, do_hard_obj
        self.do_hard_rpn_reg, self.do_hard_roi_reg = do_hard_rpn_reg, do_hard_roi_reg
        self.do_cls_dst, self.do_obj_dst = do_cls_dst, do_obj_dst
        self.do_rpn_reg_dst, self.do_roih_reg_dst = do_rpn_reg_dst, do_roih_reg_dst
        self.cls_temperature, self.obj_temperature = cls_temperature, obj_temperature
        self.cls_loss_type = cls_loss_type
        self.threshold = pseudo_label_threshold
        self.io = {}
        self._replacement = None
        self.seed = random.randint(0, 2 ** 32 - 1)  # aldi/helpers.py:19-23
        self._register_hooks()

    def _register_hooks(self):  # aldi/distill.py:115-138
        io = self.io
        s, t = self.student, self.teacher
        s.proposal_generator.register_forward_hook(lambda m, i, o: io.__setitem__("s_rpn", o))
        s.proposal_generator.rpn_head.register_forward_hook(lambda m, i, o: io.__setitem__("s_rpn_head", o))
        s.roi_heads.box_predictor.register_forward_hook(lambda m, i, o: io.__setitem__("s_boxpred", o))
        t.backbone.register_forward_hook(lambda m, i, o: io.__setitem__("t_backbone", o))
        t.proposal_generator.rpn_head.register_forward_hook(lambda m, i, o: io.__setitem__("t_rpn_head", o))
        t.roi_heads.box_predictor.register_forward_hook(lambda m, i, o: io.__setitem__("t_boxpred", o))
        t.proposal_generator.anchor_generator.register_forward_hook(lambda m, i, o: io.__setitem__("t_anchors", o))

        def seed_hook(module, args):  # aldi/helpers.py:25-26 (T3)
            torch.manual_seed(self.seed)

        t.roi_heads.register_forward_pre_hook(seed_hook)
        s.roi_heads.register_forward_pre_hook(seed_hook)

        def replace_once(module, args):  # aldi/helpers.py:36-42
            if self._replacement is not None and module.training:
                images, features, proposals, gt_instances = args
                ret = (images, features, self._replacement, gt_instances)
                self._replacement = None
                return ret
            return None

        t.roi_heads.register_forward_pre_hook(replace_once)

    def distill_enabled(self):
        return any([self.do_hard_cls, self.do_hard_obj, self.do_hard_rpn_reg, self.do_hard_roi_reg, self.do_cls_dst,
                    self.do_obj_dst, self.do_rpn_reg_dst, self.do_roih_reg_dst])

    def _distill_forward(self, teacher_inputs, student_inputs):  # aldi/distill.py:144-168
        pseudo_label_inplace(self.teacher, teacher_inputs, student_inputs, self.threshold)
        self.seed = random.randint(0, 2 ** 32 - 1)
        was_eval = not self.teacher.training
        if was_eval:
            self.teacher.train()
        standard_losses = self.student(student_inputs)
        student_proposals, _ = self.io["s_rpn"]
        self._replacement = student_proposals
        with torch.no_grad():
            self.teacher(teacher_inputs)
        if was_eval:
            self.teacher.eval()
        return standard_losses

    def __call__(self, teacher_inputs, student_inputs):  # aldi/distill.py:170-191
        losses = {}
        hard = self._distill_forward(teacher_inputs, student_inputs)
        keep = {"loss_cls": self.do_hard_cls, "loss_rpn_cls": self.do_hard_obj, "loss_rpn_loc": self.do_hard_rpn_reg,
                "loss_box_reg": self.do_hard_roi_reg}
        for k, v in hard.items():
            losses[k] = v if keep.get(k, False) else v * 0.0  # T5
        losses.update(self.get_rpn_losses(teacher_inputs))
        losses.update(self.get_roih_losses())
        return losses

    def get_rpn_losses(self, teacher_inputs):  # aldi/distill.py:193-229 (traps T1, T2)
        losses = {}
        s_logits, s_deltas = self.io["s_rpn_head"]
        t_logits, t_deltas = self.io["t_rpn_head"]
        rpn = self.teacher.proposal_generator
        d2.SAMPLE_CTX["site_override"] = "rpn_distill"
        labels = torch.stack(rpn.label_and_sample_anchors(
            self.io["t_anchors"], [i["instances"].to(self.teacher.device) for i in teacher_inputs])[0])
        d2.SAMPLE_CTX["site_override"] = None
        self.io["distill_labels"] = labels
        valid_mask = torch.flatten(labels >= 0)
        fg_mask = labels == 1
        t_probs = torch.sigmoid(d2.cat([torch.flatten(t) for t in t_logits]) / self.obj_temperature)
        if self.do_obj_dst:
            losses["loss_obj_bce"] = F.binary_cross_entropy_with_logits(
                d2.cat([torch.flatten(t) for t in s_logits])[valid_mask], t_probs[valid_mask], reduction="mean")
        if self.do_rpn_reg_dst:
            fg4 = torch.repeat_interleave(fg_mask, repeats=4)
            losses["loss_rpn_l1"] = d2.smooth_l1_loss(d2.cat([torch.flatten(t) for t in s_deltas])[fg4],
                                                      d2.cat([torch.flatten(t) for t in t_deltas])[fg4], beta=0.0,
                                                      reduction="mean")
        return losses

    def get_roih_losses(self):  # aldi/distill.py:231-278
        losses = {}
        s_cls, s_deltas = self.io["s_boxpred"]
        t_cls, t_deltas = self.io["t_boxpred"]
        t_probs = F.softmax(t_cls / self.cls_temperature, dim=1)
        if self.do_cls_dst:
            if self.cls_loss_type == "CE":
                losses["loss_cls_ce"] = d2.cross_entropy(s_cls, t_probs)
            elif self.cls_loss_type == "KL":
                losses["loss_cls_ce"] = F.kl_div(F.log_softmax(s_cls, dim=1),
                                                 F.log_softmax(t_cls / self.cls_temperature, dim=1),
                                                 reduction="batchmean", log_target=True)
            else:
                raise ValueError("cls_loss_type must be one of {CE, KL}")
        if self.do_roih_reg_dst:
            bg_idx = t_cls.shape[1] - 1
            fg_cls = torch.argmax(t_cls, dim=1)
            fg = fg_cls != bg_idx
            t_fg = t_deltas.view(-1, bg_idx, 4)[fg, fg_cls[fg], :]
            s_fg = s_deltas.view(-1, bg_idx, 4)[fg, fg_cls[fg], :]
            losses["loss_roih_l1"] = d2.smooth_l1_loss(s_fg, t_fg, beta=0.0, reduction="sum") / t_cls.shape[0]
        return losses


class NullDistiller:  # aldi/distill.py:44-57
    def __call__(self, t, s):
        return {}

    def distill_enabled(self):
        return False


# ---------------------------------------------------------------------------------------------------
# aldi/trainer.py:28-117  run_model_labeled_unlabeled (+ aldi/dropin.py:94-121 SimpleTrainer.run_step)
# ---------------------------------------------------------------------------------------------------
def run_model_labeled_unlabeled(model, distiller, data, model_batch_size, backward_at_end, do_backward):
    labeled_weak, labeled_strong, unlabeled_weak, unlabeled_strong = data
    do_weak = labeled_weak is not None
    do_strong = labeled_strong is not None
    do_align = any(getattr(model, a, None) is not None for a in ["img_align", "ins_align"])
    do_distill = distiller.distill_enabled()
    total_batch_size = sum(len(s or []) for s in [labeled_weak, labeled_strong, unlabeled_weak])
    num_grad_accum_steps = total_batch_size // model_batch_size  # T9
    loss_dict = {}

    def add_to_loss_dict(losses, suffix, cond):
        for k, v in losses.items():
            if cond(k):
                v /= num_grad_accum_steps  # in place (aldi/trainer.py:70), T6
                if not backward_at_end:
                    v = v.detach()
                loss_dict[f"{k}_{suffix}"] = loss_dict.get(f"{k}_{suffix}", 0) + v

    def maybe_do_backward(losses, cond):
        if not backward_at_end:
            losses = {k: v * 0 if not cond(k) else v for k, v in losses.items()}
            do_backward(sum(losses.values()) / num_grad_accum_steps)

    passes = [0]

    def do_training_step(d, name, cond, **kw):
        for i in range(0, len(d), model_batch_size):
            d2.SAMPLE_CTX["pass"] = passes[0]
            passes[0] += 1
            loss = model(d[i:i + model_batch_size], **kw)
            maybe_do_backward(loss, cond)
            add_to_loss_dict(loss, name, cond)

    if do_weak:
        do_training_step(labeled_weak, "source_weak", lambda k: do_weak or (do_align and "_da_" in k), do_align=do_align)
    if do_strong:
        do_training_step(labeled_strong, "source_strong", lambda k: do_strong or (do_align and "_da_" in k),
                         do_align=do_align)
    if do_align:
        do_training_step(unlabeled_weak, "target_weak", lambda k: "_da_" in k, labeled=False, do_align=True)
    if do_distill:
        assert len(unlabeled_weak) == len(unlabeled_strong)
        for i in range(0, len(unlabeled_weak), model_batch_size):
            d2.SAMPLE_CTX["pass"] = 100 + passes[0]
            passes[0] += 1
            dl = distiller(unlabeled_weak[i:i + model_batch_size], unlabeled_strong[i:i + model_batch_size])
            maybe_do_backward(dl, lambda k: k != "_")
            add_to_loss_dict(dl, "distill", lambda k: k != "_")
    return loss_dict


class OracleTrainer:
    """One process, fp32, SGD: ALDITrainer.before_step + SimpleTrainer.run_step restated
    (aldi/trainer.py:242-246, aldi/dropin.py:94-121)."""

    def __init__(self, student, *, ema_alpha=0.9996, ema_start_iter=0, distill_kwargs=None, ims_per_gpu=2,
                 backward_at_end=False, lr=0.06, momentum=0.9, weight_decay=1e-4):
        self.model = student
        self.ema = EMA(student, ema_alpha, ema_start_iter)
        self.distiller = ALDIDistiller(self.ema.model, student, **distill_kwargs) if distill_kwargs else NullDistiller()
        self.ims_per_gpu = ims_per_gpu
        self.backward_at_end = backward_at_end
        params = [p for p in student.parameters() if p.requires_grad]
        self.optimizer = torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.iter = 0

    def step(self, data):
        self.ema.update_weights(self.model, self.iter)  # before_step
        self.model.train()
        with d2.EventStorage(self.iter) as storage:
            if not self.backward_at_end:
                self.optimizer.zero_grad()
            loss_dict = run_model_labeled_unlabeled(self.model, self.distiller, data, self.ims_per_gpu,
                                                    self.backward_at_end, lambda l: l.backward())
            if self.backward_at_end:
                self.optimizer.zero_grad()
                sum(loss_dict.values()).backward()
            self.optimizer.step()
        self.iter += 1
        self.last_storage = storage
        return {k: float(v) for k, v in loss_dict.items()}
