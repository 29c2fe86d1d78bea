"""The reference's plug-in surface for this path, selected by the same cfg strings:

    build_aldi(cfg)                              aldi/model.py:12-34     META_ARCH x ALIGN_MIXIN x DISTILL_MIXIN registries
    EMA(model, alpha, start_iter)                aldi/ema.py:8-60
    build_distiller(cfg, teacher, student)       aldi/distill.py:36-41   DISTILLER_REGISTRY: Distiller | HardDistiller | ALDIDistiller

The objects are FACADES over one `B200TrainStep` (the engine that owns the flat student / teacher / gradient buffers):
`model(batched_inputs, labeled=, do_align=)` runs the student forward of one micro-batch on the device and returns the
reference's loss dict as 0-d tensors that are attached to the autograd graph through `_LossBridge`; when the caller
back-propagates (`do_backward(sum(losses) / num_grad_accum_steps)`, aldi/trainer.py:75-79) the bridge receives
d(total)/d(loss_k) -- i.e. 1/num_grad_accum_steps, or 0 for the keys the trainer masked -- and runs the explicit device
backward with exactly those weights.  So the reference's own loop (`run_model_labeled_unlabeled`, aldi/trainer.py:28-117)
drives these objects unchanged; tests/test_gpu_facade.py does that with the pinned restatement of the loop.  The
benchmarked path is `B200TrainStep.run_model` (planned, fused, graph-replayed); both run the same kernels.
"""
import torch

from .registry import ALIGN_MIXIN_REGISTRY, DISTILL_MIXIN_REGISTRY, DISTILLER_REGISTRY, META_ARCH_REGISTRY
from .train_step import B200TrainStep


def _to_plain(batch):
    """Detectron2 dataset dicts -> the plain-tensor dicts the engine stages (image uint8 CHW BGR, gt boxes / classes)."""
    out = []
    for d in batch:
        if "boxes" in d or "instances" not in d:
            out.append(d)
            continue
        e = {"image": d["image"], "height": d["image"].shape[1], "width": d["image"].shape[2]}
        inst = d["instances"]
        if getattr(inst, "has", lambda k: False)("gt_boxes"):
            e["boxes"], e["classes"] = inst.gt_boxes.tensor, inst.gt_classes
        out.append(e)
    return out


def _with_empty_gt(batch):
    out = []
    for d in batch:
        e = dict(d)
        e.setdefault("boxes", torch.zeros(0, 4))
        e.setdefault("classes", torch.zeros(0, dtype=torch.int64))
        out.append(e)
    return out


class _LossBridge(torch.autograd.Function):
    """values -> one tensor in the autograd graph; backward hands the per-loss weights to the engine."""

    @staticmethod
    def forward(ctx, anchor, values, keys, finish):
        ctx.keys, ctx.finish = keys, finish
        return values.clone()

    @staticmethod
    def backward(ctx, grad):
        w = grad.detach().cpu().tolist()            # one D2H read per micro-batch backward (facade path only)
        ctx.finish(dict(zip(ctx.keys, w)))
        return None, None, None, None


def _bridge(keys, values, finish):
    anchor = torch.zeros((), device=values.device, requires_grad=True)
    out = _LossBridge.apply(anchor, values, keys, finish)
    # `* 1.0`: plain tensors of their own, so the trainer's in-place `v /= num_grad_accum_steps` (aldi/trainer.py:70) is legal
    return {k: out[i] * 1.0 for i, k in enumerate(keys)}


# ---------------------------------------------------------------------------------------------------------------------
@META_ARCH_REGISTRY.register()
class GeneralizedRCNN:
    """Facade with the call surface of detectron2's GeneralizedRCNN that the ALDI path uses: `forward` (training loss
    dict), `inference(batched_inputs, do_postprocess)`, `state_dict` / `load_state_dict` with Detectron2 key names,
    `device`, `training` / `train()` / `eval()`."""

    def __init__(self, cfg=None, *, engine=None, which="student", state_dict=None, device=None, process_group=None,
                 dtype=None, **kwargs):
        if engine is None:
            from . import arch
            from .config import step_config_from_cfg
            scfg = step_config_from_cfg(cfg, dtype=dtype)
            if state_dict is None:
                from .train_step import synthetic_state_dict_for
                state_dict = synthetic_state_dict_for(scfg)
            dev = device or (cfg.MODEL.DEVICE if cfg.MODEL.DEVICE != "cuda" else "cuda:%d" % torch.cuda.current_device())
            engine = B200TrainStep(scfg, state_dict, device=dev, process_group=process_group)
            engine.debug = None
        self.engine, self.which, self.training = engine, which, True
        # aldi/align.py:41-42: the trainer asks `getattr(model, "img_align")` to decide on alignment passes
        self.img_align = "img_align" if engine.cfg.img_da_enabled else None
        self.ins_align = "ins_align" if engine.cfg.ins_da_enabled else None

    @property
    def device(self):
        return self.engine.device

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        assert torch.device(device).type == self.device.type, "the engine lives on the device it was built on"
        return self

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, batched_inputs, labeled=True, do_align=False):
        if not self.training:
            return self.inference(batched_inputs)
        assert self.which == "student", "only the student is trained; the teacher is the EMA of it (aldi/ema.py)"
        data = _to_plain(batched_inputs)
        if labeled:
            data = _with_empty_gt(data)        # an image without annotations is an empty ground truth
        keys, vals, finish = self.engine.facade_forward(data, labeled=labeled, do_align=do_align)
        out = _bridge(keys, vals, finish)
        if not do_align and (self.img_align or self.ins_align):
            out["_da"] = torch.zeros((), device=vals.device)       # aldi/align.py:91-100 placeholder
        return out

    def inference(self, batched_inputs, do_postprocess=True):
        return self.engine.inference(_to_plain(batched_inputs), which=self.which, do_postprocess=do_postprocess)

    def state_dict(self):
        return self.engine.state_dict(self.which)

    def load_state_dict(self, sd, strict=True):
        return self.engine.load_state_dict(sd, which=self.which, strict=strict)


@ALIGN_MIXIN_REGISTRY.register()
class AlignMixin(GeneralizedRCNN):
    """aldi/align.py:17-101.  The discriminators live in the engine's flat buffers (`img_align.*`, `ins_align.*` keys)
    and are switched on by DOMAIN_ADAPT.ALIGN.* through `step_config_from_cfg`; nothing to add on the facade."""


@DISTILL_MIXIN_REGISTRY.register()
class DistillMixin(GeneralizedRCNN):
    """aldi/distill.py:281-285: no modification of the module is needed."""


def build_aldi(cfg, **kwargs):
    """aldi/model.py:12-34: `class ALDI(align_mixin, distill_mixin, base_cls)` from the three registries named in cfg."""
    base_cls = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)
    align_mixin = ALIGN_MIXIN_REGISTRY.get(cfg.DOMAIN_ADAPT.ALIGN.MIXIN_NAME)
    distill_mixin = DISTILL_MIXIN_REGISTRY.get(cfg.DOMAIN_ADAPT.DISTILL.MIXIN_NAME)

    class ALDI(align_mixin, distill_mixin, base_cls):
        def forward(self, batched_inputs, labeled=True, do_align=False):
            return super(ALDI, self).forward(batched_inputs, do_align=do_align, labeled=labeled)

    return ALDI(cfg, **kwargs)


# ---------------------------------------------------------------------------------------------------------------------
class EMA:
    """aldi/ema.py:8-60 over the engine's flat teacher buffer: `.model` is the teacher facade of the SAME engine."""

    def __init__(self, model, alpha, start_iter=0):
        self.engine = model.engine
        self.model = type(model)(engine=model.engine, which="teacher")
        self.alpha, self.start_iter = alpha, start_iter
        self.engine.cfg.ema_alpha, self.engine.cfg.ema_start_iter = alpha, start_iter

    def update_weights(self, model, iter):
        assert model.engine is self.engine
        self.engine.ema_update(iter)

    def inference(self, batched_inputs):
        return self.model.inference(batched_inputs)

    def state_dict(self):
        return {"model." + k: v for k, v in self.model.state_dict().items()}


# ---------------------------------------------------------------------------------------------------------------------
def build_distiller(cfg, teacher, student):
    """aldi/distill.py:36-41."""
    return DISTILLER_REGISTRY.get(cfg.DOMAIN_ADAPT.DISTILL.DISTILLER_NAME).from_config(cfg, teacher, student)


@DISTILLER_REGISTRY.register()
class Distiller:
    """This Distiller does nothing (aldi/distill.py:44-57)."""

    def __init__(self, teacher, student):
        self.teacher, self.student = teacher, student

    @classmethod
    def from_config(cls, cfg, teacher, student):
        return Distiller(teacher, student)

    def __call__(self, teacher_batched_inputs, student_batched_inputs):
        return {}

    def distill_enabled(self):
        return False

    # flags the engine's planned step (`B200TrainStep.run_model`) runs with when this distiller is configured
    def engine_flags(self):
        return dict(do_hard_cls=False, do_hard_obj=False, do_hard_rpn_reg=False, do_hard_roi_reg=False, do_cls_dst=False,
                    do_obj_dst=False, do_rpn_reg_dst=False, do_roih_reg_dst=False)

    def _apply_flags(self):
        eng = getattr(self.student, "engine", None)
        if eng is not None:
            for k, v in self.engine_flags().items():
                setattr(eng.cfg, k, v)


@DISTILLER_REGISTRY.register()
class HardDistiller(Distiller):
    """Hard pseudo-label self-distillation only (aldi/distill.py:60-85): the student's standard losses against the
    thresholded teacher detections; `distill_enabled` looks at the four HARD_* flags alone."""

    def __init__(self, teacher, student, do_hard_cls=False, do_hard_obj=False, do_hard_rpn_reg=False, do_hard_roi_reg=False,
                 pseudo_label_threshold=0.8):
        super().__init__(teacher, student)
        self.do_hard_cls, self.do_hard_obj = do_hard_cls, do_hard_obj
        self.do_hard_rpn_reg, self.do_hard_roi_reg = do_hard_rpn_reg, do_hard_roi_reg
        self.pseudo_label_threshold = pseudo_label_threshold
        self._apply_flags()

    @classmethod
    def from_config(cls, cfg, teacher, student):
        D = cfg.DOMAIN_ADAPT.DISTILL
        return HardDistiller(teacher, student, do_hard_cls=D.HARD_ROIH_CLS_ENABLED, do_hard_obj=D.HARD_OBJ_ENABLED,
                             do_hard_rpn_reg=D.HARD_RPN_REG_ENABLED, do_hard_roi_reg=D.HARD_ROIH_REG_ENABLED,
                             pseudo_label_threshold=cfg.DOMAIN_ADAPT.TEACHER.THRESHOLD)

    def engine_flags(self):
        # HardDistiller.__call__ returns the student's standard losses UNMASKED (aldi/distill.py:78-81): every hard loss
        # of the pass trains, whichever of the flags switched the distiller on
        on = self.distill_enabled()
        return dict(super().engine_flags(), do_hard_cls=on, do_hard_obj=on, do_hard_rpn_reg=on, do_hard_roi_reg=on)

    def distill_enabled(self):
        return any([self.do_hard_cls, self.do_hard_obj, self.do_hard_rpn_reg, self.do_hard_roi_reg])

    def _call_engine(self, teacher_batched_inputs, student_batched_inputs):
        eng = self.student.engine
        assert self.teacher.engine is eng, "teacher and student must be the two sides of one engine"
        self._apply_flags()
        eng.cfg.pseudo_threshold = self.pseudo_label_threshold
        keys, vals, finish = eng.facade_distill(_to_plain(teacher_batched_inputs), _to_plain(student_batched_inputs))
        return _bridge(keys, vals, finish)

    def __call__(self, teacher_batched_inputs, student_batched_inputs):
        return self._call_engine(teacher_batched_inputs, student_batched_inputs)


@DISTILLER_REGISTRY.register()
class ALDIDistiller(HardDistiller):
    """Hard and / or soft distillation for Faster R-CNN students and teachers (aldi/distill.py:87-278)."""

    def __init__(self, teacher, student, do_hard_cls=False, do_hard_obj=False, do_hard_rpn_reg=False, do_hard_roi_reg=False,
                 do_cls_dst=False, do_obj_dst=False, do_rpn_reg_dst=False, do_roih_reg_dst=False, cls_temperature=1.0,
                 obj_temperature=1.0, cls_loss_type="CE", pseudo_label_threshold=0.8):
        self.do_cls_dst, self.do_obj_dst = do_cls_dst, do_obj_dst
        self.do_rpn_reg_dst, self.do_roih_reg_dst = do_rpn_reg_dst, do_roih_reg_dst
        self.cls_temperature, self.obj_temperature, self.cls_loss_type = cls_temperature, obj_temperature, cls_loss_type
        super().__init__(teacher, student, do_hard_cls, do_hard_obj, do_hard_rpn_reg, do_hard_roi_reg, pseudo_label_threshold)

    @classmethod
    def from_config(cls, cfg, teacher, student):
        D = cfg.DOMAIN_ADAPT.DISTILL
        return ALDIDistiller(teacher, student, do_hard_cls=D.HARD_ROIH_CLS_ENABLED, do_hard_obj=D.HARD_OBJ_ENABLED,
                             do_hard_rpn_reg=D.HARD_RPN_REG_ENABLED, do_hard_roi_reg=D.HARD_ROIH_REG_ENABLED,
                             do_cls_dst=D.ROIH_CLS_ENABLED, do_obj_dst=D.OBJ_ENABLED, do_rpn_reg_dst=D.RPN_REG_ENABLED,
                             do_roih_reg_dst=D.ROIH_REG_ENABLED, cls_temperature=D.CLS_TMP, obj_temperature=D.OBJ_TMP,
                             cls_loss_type=cfg.DOMAIN_ADAPT.CLS_LOSS_TYPE,
                             pseudo_label_threshold=cfg.DOMAIN_ADAPT.TEACHER.THRESHOLD)

    def engine_flags(self):
        return dict(do_hard_cls=self.do_hard_cls, do_hard_obj=self.do_hard_obj, do_hard_rpn_reg=self.do_hard_rpn_reg,
                    do_hard_roi_reg=self.do_hard_roi_reg, do_cls_dst=self.do_cls_dst, do_obj_dst=self.do_obj_dst,
                    do_rpn_reg_dst=self.do_rpn_reg_dst, do_roih_reg_dst=self.do_roih_reg_dst)

    def distill_enabled(self):
        return any(self.engine_flags().values())

    def __call__(self, teacher_batched_inputs, student_batched_inputs):
        eng = self.student.engine
        eng.cfg.cls_temperature, eng.cfg.obj_temperature = self.cls_temperature, self.obj_temperature
        eng.cfg.cls_loss_type = self.cls_loss_type
        out = self._call_engine(teacher_batched_inputs, student_batched_inputs)
        # aldi/distill.py:181-186 (T5): hard losses that are switched off stay in the dict, multiplied by 0.0
        keep = {"loss_cls": self.do_hard_cls, "loss_rpn_cls": self.do_hard_obj, "loss_rpn_loc": self.do_hard_rpn_reg,
                "loss_box_reg": self.do_hard_roi_reg}
        return {k: (v if keep.get(k, True) else v * 0.0) for k, v in out.items()}
