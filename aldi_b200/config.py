"""Config surface of the hot path: a yacs-style CfgNode (attribute dict, `_BASE_` YAML inheritance,
merge_from_list) so the reference's shipped YAMLs (configs/**.yaml) load unchanged, the Detectron2 defaults
the step reads, and `add_aldi_config` with the reference's keys and defaults (aldi/config.py:7-100).
"""
import ast
import copy
import os

import yaml

from .train_step import StepConfig


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        if self.__dict__["_frozen"]:
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(k, v))
        self[k] = v

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        c = CfgNode()
        for k, v in self.items():
            dict.__setitem__(c, k, copy.deepcopy(v, memo))
        return c

    # ---- merging ---------------------------------------------------------------------------------------
    @staticmethod
    def load_yaml_with_base(filename):
        with open(filename) as fh:
            cfg = yaml.safe_load(fh) or {}
        base = cfg.pop("_BASE_", None)
        if base is not None:
            if not os.path.isabs(base):
                base = os.path.join(os.path.dirname(filename), base)
            merged = CfgNode.load_yaml_with_base(base)
            _merge_dict(merged, cfg)
            return merged
        return cfg

    def merge_from_file(self, filename, allow_unsafe=False):
        loaded = CfgNode.load_yaml_with_base(filename)
        loaded.pop("VERSION", None)
        _merge_into(self, loaded, [])

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0, "Override list has odd length: {}".format(opts)
        for full_key, v in zip(opts[0::2], opts[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                if k not in node:
                    raise KeyError("Non-existent config key: {}".format(full_key))
                node = node[k]
            if keys[-1] not in node:
                raise KeyError("Non-existent config key: {}".format(full_key))
            node[keys[-1]] = _coerce(_decode(v), node[keys[-1]], full_key)


def _decode(v):
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _coerce(new, old, key):
    if old is None or new is None or type(new) == type(old):
        return new
    for a, b in ((tuple, list), (list, tuple)):
        if isinstance(old, a) and isinstance(new, b):
            return a(new)
    if isinstance(old, float) and isinstance(new, int):
        return float(new)
    if isinstance(old, (tuple, list)) and isinstance(new, str):  # yaml "(1,1)" tuples
        return type(old)(_decode(new))
    raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(old), type(new), key))


def _merge_dict(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_dict(dst[k], v)
        else:
            dst[k] = v


def _merge_into(node, src, path):
    for k, v in src.items():
        full = ".".join(path + [k])
        if k not in node:
            raise KeyError("Non-existent config key: {}".format(full))
        if isinstance(v, dict):
            _merge_into(node[k], v, path + [k])
        else:
            node[k] = _coerce(_decode(v), node[k], full)


def get_cfg():
    """The Detectron2 defaults read on the ALDI hot path (detectron2/config/defaults.py) plus every key the
    shipped R-CNN YAML chain (configs/detectron2/Base-RCNN-FPN.yaml -> configs/Base-RCNN-FPN.yaml ->
    configs/cityscapes/*.yaml) sets."""
    C = CfgNode
    c = C()
    c.VERSION = 2
    c.MODEL = C(dict(
        META_ARCHITECTURE="GeneralizedRCNN", DEVICE="cuda", WEIGHTS="", MASK_ON=False, KEYPOINT_ON=False,
        PIXEL_MEAN=[103.530, 116.280, 123.675], PIXEL_STD=[1.0, 1.0, 1.0],
        BACKBONE=dict(NAME="build_resnet_fpn_backbone", FREEZE_AT=2),
        RESNETS=dict(DEPTH=50, OUT_FEATURES=["res2", "res3", "res4", "res5"], STRIDE_IN_1X1=True, NORM="FrozenBN",
                     NUM_GROUPS=1, WIDTH_PER_GROUP=64, RES2_OUT_CHANNELS=256, STEM_OUT_CHANNELS=64),
        FPN=dict(IN_FEATURES=["res2", "res3", "res4", "res5"], OUT_CHANNELS=256, NORM="", FUSE_TYPE="sum"),
        ANCHOR_GENERATOR=dict(NAME="DefaultAnchorGenerator", SIZES=[[32], [64], [128], [256], [512]],
                              ASPECT_RATIOS=[[0.5, 1.0, 2.0]], OFFSET=0.0),
        PROPOSAL_GENERATOR=dict(NAME="RPN", MIN_SIZE=0),
        RPN=dict(HEAD_NAME="StandardRPNHead", IN_FEATURES=["p2", "p3", "p4", "p5", "p6"], BOUNDARY_THRESH=-1,
                 IOU_THRESHOLDS=[0.3, 0.7], IOU_LABELS=[0, -1, 1], BATCH_SIZE_PER_IMAGE=256, POSITIVE_FRACTION=0.5,
                 BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0, BBOX_REG_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
                 SMOOTH_L1_BETA=0.0, LOSS_WEIGHT=1.0, PRE_NMS_TOPK_TRAIN=12000, PRE_NMS_TOPK_TEST=6000,
                 POST_NMS_TOPK_TRAIN=2000, POST_NMS_TOPK_TEST=1000, NMS_THRESH=0.7, CONV_DIMS=[-1]),
        ROI_HEADS=dict(NAME="StandardROIHeads", NUM_CLASSES=80, IN_FEATURES=["p2", "p3", "p4", "p5"],
                       IOU_THRESHOLDS=[0.5], IOU_LABELS=[0, 1], BATCH_SIZE_PER_IMAGE=512, POSITIVE_FRACTION=0.25,
                       SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5, PROPOSAL_APPEND_GT=True),
        ROI_BOX_HEAD=dict(NAME="FastRCNNConvFCHead", BBOX_REG_LOSS_TYPE="smooth_l1", BBOX_REG_LOSS_WEIGHT=1.0,
                          BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), SMOOTH_L1_BETA=0.0, POOLER_RESOLUTION=7,
                          POOLER_SAMPLING_RATIO=0, POOLER_TYPE="ROIAlignV2", NUM_FC=2, FC_DIM=1024, NUM_CONV=0,
                          CONV_DIM=256, NORM="", CLS_AGNOSTIC_BBOX_REG=False, TRAIN_ON_PRED_BOXES=False),
        ROI_MASK_HEAD=dict(NAME="MaskRCNNConvUpsampleHead", NUM_CONV=4, POOLER_RESOLUTION=14),
    ))
    c.INPUT = C(dict(MIN_SIZE_TRAIN=(800,), MIN_SIZE_TRAIN_SAMPLING="choice", MAX_SIZE_TRAIN=1333, MIN_SIZE_TEST=800,
                     MAX_SIZE_TEST=1333, FORMAT="BGR", RANDOM_FLIP="horizontal"))
    c.DATASETS = C(dict(TRAIN=(), TEST=()))
    c.DATALOADER = C(dict(NUM_WORKERS=4, FILTER_EMPTY_ANNOTATIONS=True))
    c.SOLVER = C(dict(LR_SCHEDULER_NAME="WarmupMultiStepLR", MAX_ITER=40000, BASE_LR=0.001, MOMENTUM=0.9, NESTEROV=False,
                      WEIGHT_DECAY=0.0001, WEIGHT_DECAY_NORM=0.0, GAMMA=0.1, STEPS=(30000,), WARMUP_FACTOR=1.0 / 1000,
                      WARMUP_ITERS=1000, WARMUP_METHOD="linear", CHECKPOINT_PERIOD=5000, IMS_PER_BATCH=16,
                      BIAS_LR_FACTOR=1.0, WEIGHT_DECAY_BIAS=None, AMP=dict(ENABLED=False)))
    c.TEST = C(dict(EVAL_PERIOD=0, DETECTIONS_PER_IMAGE=100))
    c.OUTPUT_DIR = "./output"
    c.SEED = -1
    c.VIS_PERIOD = 0
    return c


def add_aldi_config(cfg):
    """Same keys and defaults as the reference's aldi/config.py:7-100 (everything off unless enabled)."""
    C = CfgNode
    cfg.DATASETS.UNLABELED = tuple()
    cfg.DATASETS.BATCH_CONTENTS = ("labeled_weak",)
    cfg.DATASETS.BATCH_RATIOS = (1,)
    cfg.AUG = C(dict(WEAK_INCLUDES_MULTISCALE=True, LABELED_INCLUDE_RANDOM_ERASING=True,
                     UNLABELED_INCLUDE_RANDOM_ERASING=True, LABELED_MIC_AUG=False, UNLABELED_MIC_AUG=False,
                     MIC_RATIO=0.5, MIC_BLOCK_SIZE=32))
    cfg.EMA = C(dict(ENABLED=False, ALPHA=0.9996, LOAD_FROM_EMA_ON_START=True, START_ITER=0))
    cfg.DOMAIN_ADAPT = C(dict(
        ALIGN=dict(MIXIN_NAME="AlignMixin", IMG_DA_ENABLED=False, IMG_DA_LAYER="p2", IMG_DA_WEIGHT=0.01,
                   IMG_DA_INPUT_DIM=256, IMG_DA_HIDDEN_DIMS=[256], INS_DA_ENABLED=False, INS_DA_WEIGHT=0.01,
                   INS_DA_INPUT_DIM=1024, INS_DA_HIDDEN_DIMS=[1024]),
        DISTILL=dict(DISTILLER_NAME="ALDIDistiller", MIXIN_NAME="DistillMixin", HARD_ROIH_CLS_ENABLED=False,
                     HARD_ROIH_REG_ENABLED=False, HARD_OBJ_ENABLED=False, HARD_RPN_REG_ENABLED=False,
                     ROIH_CLS_ENABLED=False, ROIH_REG_ENABLED=False, OBJ_ENABLED=False, RPN_REG_ENABLED=False,
                     CLS_TMP=1.0, OBJ_TMP=1.0),
        CLS_LOSS_TYPE="CE",
        TEACHER=dict(ENABLED=False, THRESHOLD=0.8)))
    cfg.VIT = C(dict(USE_ACT_CHECKPOINT=True))
    cfg.SOLVER.IMS_PER_GPU = 2
    cfg.SOLVER.BACKWARD_AT_END = True
    cfg.SOLVER.OPTIMIZER = "SGD"
    cfg.MODEL.CONVNEXT = C(dict(DEPTHS=[3, 3, 9, 3], DIMS=[96, 192, 384, 768], DROP_PATH_RATE=0.2,
                                LAYER_SCALE_INIT_VALUE=1e-6, OUT_FEATURES=[0, 1, 2, 3]))
    cfg.SOLVER.WEIGHT_DECAY_RATE = 0.95
    return cfg


def step_config_from_cfg(cfg, dtype=None):
    """cfg (reference key names) -> the StepConfig the B200 step consumes."""
    D = cfg.DOMAIN_ADAPT.DISTILL
    backbones = {"build_resnet_fpn_backbone": "resnet50", "build_convnext_fpn_backbone": "convnext",
                 "build_vitdet_b_backbone": "vitdet_b", "build_vitdet_l_backbone": "vitdet_l"}
    if cfg.MODEL.META_ARCHITECTURE != "GeneralizedRCNN" or cfg.MODEL.BACKBONE.NAME not in backbones:
        raise NotImplementedError("Faster R-CNN on ResNet-50-FPN (BASELINE configs[0-1]), ViTDet-B/L (configs[2]) and "
                                  "ConvNeXt-FPN (configs[4]) are built; got %s / %s"
                                  % (cfg.MODEL.META_ARCHITECTURE, cfg.MODEL.BACKBONE.NAME))
    backbone = backbones[cfg.MODEL.BACKBONE.NAME]
    B, R = cfg.MODEL.ROI_BOX_HEAD, cfg.MODEL.RPN
    if backbone.startswith("vitdet"):
        # configs/Base-RCNN-VitDetB.yaml:7-14: two-conv RPN head, four 3x3 convs + "LN" and one FC in the box head
        if list(R.CONV_DIMS) != [-1, -1] or (B.NUM_CONV, B.CONV_DIM, B.NORM, B.NUM_FC, B.FC_DIM) != (4, 256, "LN", 1, 1024):
            raise NotImplementedError("ViTDet heads: MODEL.RPN.CONV_DIMS [-1, -1] and ROI_BOX_HEAD NUM_CONV 4 / CONV_DIM 256 / "
                                      "NORM LN / NUM_FC 1 / FC_DIM 1024 (Base-RCNN-VitDetB.yaml) are built")
    elif list(R.CONV_DIMS) != [-1] or (B.NUM_CONV, B.NUM_FC, B.FC_DIM) != (0, 2, 1024):
        raise NotImplementedError("FPN detectors: the standard RPN head (CONV_DIMS [-1]) and the 2-FC box head are built")
    optimizer = (cfg.SOLVER.OPTIMIZER or "SGD").upper()
    if optimizer not in ("SGD", "ADAMW"):                       # aldi/trainer.py:207-208
        raise ValueError("Unsupported optimizer/backbone combination {} {}.".format(cfg.SOLVER.OPTIMIZER,
                                                                                     cfg.MODEL.BACKBONE.NAME))
    extra = {}
    base_lr, weight_decay = cfg.SOLVER.BASE_LR, cfg.SOLVER.WEIGHT_DECAY
    if optimizer == "ADAMW":
        # aldi/trainer.py:205-206 -> get_adamw_optim(model) with no overrides (aldi/backbone.py:66-84): the optimizer is
        # detectron2's configs/common/optim.py AdamW LazyConfig as it stands -- lr 1e-4, betas (0.9, 0.999), weight decay
        # 0.1 -- and SOLVER.BASE_LR / SOLVER.WEIGHT_DECAY are never read on this branch (the ConvNeXt yaml's BASE_LR 0.06
        # would diverge under AdamW).  The LR scheduler multiplies that 1e-4 (detectron2 LRMultiplier).  Its
        # weight_decay_norm=0.0 exempts torch.nn norm modules only; ConvNeXt's LayerNorm is the reference's own
        # nn.Module (aldi/backbone.py:331) and the FPN has no norm, so the decay is uniform over every parameter.
        base_lr, weight_decay = 1e-4, 0.1
    if backbone == "convnext":
        extra = dict(convnext_depths=tuple(cfg.MODEL.CONVNEXT.DEPTHS), convnext_dims=tuple(cfg.MODEL.CONVNEXT.DIMS),
                     convnext_drop_path=cfg.MODEL.CONVNEXT.DROP_PATH_RATE)
    if max(cfg.MODEL.RPN.PRE_NMS_TOPK_TRAIN, cfg.MODEL.RPN.PRE_NMS_TOPK_TEST) > 2048:
        raise NotImplementedError("MODEL.RPN.PRE_NMS_TOPK_TRAIN/TEST up to 2048 per level are supported (the ALDI configs "
                                  "use 2000 / 1000, configs/detectron2/Base-RCNN-FPN.yaml:14-15); got %d / %d"
                                  % (cfg.MODEL.RPN.PRE_NMS_TOPK_TRAIN, cfg.MODEL.RPN.PRE_NMS_TOPK_TEST))
    A = cfg.DOMAIN_ADAPT.ALIGN
    return StepConfig(
        backbone=backbone, optimizer=optimizer, pixel_mean=tuple(cfg.MODEL.PIXEL_MEAN), pixel_std=tuple(cfg.MODEL.PIXEL_STD),
        anchor_sizes=tuple(tuple(s) for s in cfg.MODEL.ANCHOR_GENERATOR.SIZES), **extra,
        img_da_enabled=A.IMG_DA_ENABLED, img_da_layer=A.IMG_DA_LAYER, img_da_weight=A.IMG_DA_WEIGHT,
        img_da_input_dim=A.IMG_DA_INPUT_DIM, img_da_hidden_dims=tuple(A.IMG_DA_HIDDEN_DIMS),
        ins_da_enabled=A.INS_DA_ENABLED, ins_da_weight=A.INS_DA_WEIGHT, ins_da_input_dim=A.INS_DA_INPUT_DIM,
        ins_da_hidden_dims=tuple(A.INS_DA_HIDDEN_DIMS),
        num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES, ims_per_gpu=cfg.SOLVER.IMS_PER_GPU, ema_alpha=cfg.EMA.ALPHA,
        ema_start_iter=cfg.EMA.START_ITER, pseudo_threshold=cfg.DOMAIN_ADAPT.TEACHER.THRESHOLD,
        do_hard_cls=D.HARD_ROIH_CLS_ENABLED, do_hard_obj=D.HARD_OBJ_ENABLED, do_hard_rpn_reg=D.HARD_RPN_REG_ENABLED,
        do_hard_roi_reg=D.HARD_ROIH_REG_ENABLED, do_cls_dst=D.ROIH_CLS_ENABLED, do_obj_dst=D.OBJ_ENABLED,
        do_rpn_reg_dst=D.RPN_REG_ENABLED, do_roih_reg_dst=D.ROIH_REG_ENABLED, cls_temperature=D.CLS_TMP,
        obj_temperature=D.OBJ_TMP, cls_loss_type=cfg.DOMAIN_ADAPT.CLS_LOSS_TYPE, base_lr=base_lr,
        momentum=cfg.SOLVER.MOMENTUM, weight_decay=weight_decay,
        rpn_pre_topk=(cfg.MODEL.RPN.PRE_NMS_TOPK_TRAIN, cfg.MODEL.RPN.PRE_NMS_TOPK_TEST),
        rpn_post_topk=(cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN, cfg.MODEL.RPN.POST_NMS_TOPK_TEST),
        rpn_nms_thresh=cfg.MODEL.RPN.NMS_THRESH, rpn_batch=cfg.MODEL.RPN.BATCH_SIZE_PER_IMAGE,
        rpn_pos_fraction=cfg.MODEL.RPN.POSITIVE_FRACTION, rpn_iou=tuple(cfg.MODEL.RPN.IOU_THRESHOLDS),
        roi_batch=cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE, roi_pos_fraction=cfg.MODEL.ROI_HEADS.POSITIVE_FRACTION,
        roi_iou=cfg.MODEL.ROI_HEADS.IOU_THRESHOLDS[0], test_score_thresh=cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST,
        test_nms_thresh=cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST, test_topk=cfg.TEST.DETECTIONS_PER_IMAGE,
        dtype=dtype or ("bf16" if cfg.SOLVER.AMP.ENABLED else "fp32"))
