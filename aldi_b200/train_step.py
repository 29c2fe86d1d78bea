"""The ALDI++ teacher–student training step on one B200 (one process per GPU).

`B200TrainStep` is the object the reference's trainer seams are cut at: it owns the flat
student / teacher / gradient / momentum buffers and exposes
    ema_update(iter)            <- ALDITrainer.before_step          (aldi/trainer.py:242-246, aldi/ema.py:52-57)
    run_model(data) -> losses   <- _ALDITrainer.run_model           (aldi/trainer.py:129-130, :28-117)
    optimizer_step(lr)          <- SimpleTrainer/AMPTrainer.run_step (aldi/dropin.py:121,177-178)
Semantics reproduced from the reference (SURVEY.md §8a traps): micro-batching and the 1/num_grad_accum
scaling incl. uneven micro-batches (T6, T9), zero-weighted hard losses kept in the dict (T5), the RPN
distillation index quirk (T1, in the kernel), a fresh anchor sampling on the teacher RPN with pseudo GT
(T2), one sampling seed per distillation forward shared by student and teacher RoI heads (T3), EMA over
every state_dict entry incl. FrozenBN buffers (T7).

Schedule differences that do not change results: the teacher trunk + RPN head run ONCE per micro-batch
(pass #1 eval-mode inference and pass #2 train-mode forward see identical FrozenBN features); gradients of all
micro-batches accumulate into one flat buffer that is all-reduced once per step (the reference all-reduces
per micro-batch backward; averaging is linear).
"""
import os
import random

import torch
import torch.distributed as dist

from . import lib as _l
from . import arch, ops, sampling
from .data_parallel import GradReducer
from .detector import Detector, DetectorWeights, FlatLayout

SRC_KEYS = ("loss_cls", "loss_box_reg", "loss_rpn_cls", "loss_rpn_loc")
SOFT_KEYS = ("loss_obj_bce", "loss_rpn_l1", "loss_cls_ce", "loss_roih_l1")
DA_KEYS = ("loss_da_img", "loss_da_ins")
# slots of the device-side loss vector (one D2H read per step): per source tag 4 hard + 2 alignment losses, the
# target_weak alignment pass, then the distillation pass (4 hard + 4 soft); the last slot stays 0 (the reference's
# `_da` placeholder, aldi/align.py:91-100)
LOSS_SLOTS = 32
SLOT_BASE = {"source_weak": 0, "source_strong": 6, "target_weak": 12, "distill": 14}
SLOT_ZERO = LOSS_SLOTS - 1


class StepConfig:
    """The cfg values the hot path reads (names follow aldi/config.py and detectron2 defaults)."""

    def __init__(self, **kw):
        self.num_classes = 8
        self.ims_per_gpu = 2                     # SOLVER.IMS_PER_GPU
        self.ema_alpha = 0.9996                  # EMA.ALPHA
        self.ema_start_iter = 0                  # EMA.START_ITER
        self.pseudo_threshold = 0.8              # DOMAIN_ADAPT.TEACHER.THRESHOLD
        self.do_hard_cls = self.do_hard_obj = self.do_hard_rpn_reg = self.do_hard_roi_reg = False
        self.do_cls_dst = self.do_obj_dst = self.do_rpn_reg_dst = self.do_roih_reg_dst = True
        self.cls_temperature = 1.0               # DOMAIN_ADAPT.DISTILL.CLS_TMP
        self.obj_temperature = 1.0               # DOMAIN_ADAPT.DISTILL.OBJ_TMP
        self.cls_loss_type = "CE"                # DOMAIN_ADAPT.CLS_LOSS_TYPE
        self.base_lr = 0.06                      # SOLVER.BASE_LR
        self.momentum = 0.9
        self.weight_decay = 1e-4
        self.rpn_pre_topk = (2000, 1000)         # (train, test) per level
        self.rpn_post_topk = (1000, 1000)
        self.rpn_nms_thresh = 0.7
        self.rpn_batch = 256
        self.rpn_pos_fraction = 0.5
        self.rpn_iou = (0.3, 0.7)
        self.roi_batch = 512
        self.roi_pos_fraction = 0.25
        self.roi_iou = 0.5
        self.test_score_thresh = 0.05
        self.test_nms_thresh = 0.5
        self.test_topk = 100
        # DOMAIN_ADAPT.ALIGN.* (aldi/config.py:38-49); both off in every shipped ALDI++ config
        self.img_da_enabled = False
        self.img_da_layer = "p2"
        self.img_da_weight = 0.01
        self.img_da_input_dim = 256
        self.img_da_hidden_dims = (256,)
        self.ins_da_enabled = False
        self.ins_da_weight = 0.01
        self.ins_da_input_dim = 1024
        self.ins_da_hidden_dims = (1024,)
        # bottom-up: "resnet50" (BASELINE configs[0-1]) or "convnext" (configs[4], MODEL.CONVNEXT.*, aldi/config.py:94-99)
        self.backbone = "resnet50"
        self.convnext_depths = (3, 3, 27, 3)
        self.convnext_dims = (192, 384, 768, 1536)
        self.convnext_drop_path = 0.2
        # "vitdet_b" / "vitdet_l" (BASELINE configs[2]; build_vitdet_{b,l}_backbone, aldi/backbone.py:37-64): plain ViT +
        # SimpleFeaturePyramid with the two-conv RPN head and the 4-conv LayerNorm box head of Base-RCNN-VitDetB.yaml
        self.vit_img_size = 1024                 # ViT(img_size=...) of mask_rcnn_vitdet.py: sizes the global rel-pos tables
        self.vit_overrides = None                # test seam: ViT(...) keyword overrides (embed_dim, depth, ...)
        self.anchor_sizes = None                 # MODEL.ANCHOR_GENERATOR.SIZES (None: detectron2's 32..512)
        self.pixel_mean = (103.530, 116.280, 123.675)
        self.pixel_std = (1.0, 1.0, 1.0)
        self.optimizer = "SGD"                   # SOLVER.OPTIMIZER: "SGD" | "ADAMW" (aldi/trainer.py:199-208)
        self.adamw_betas = (0.9, 0.999)
        self.adamw_eps = 1e-8
        self.dtype = "bf16"                      # "bf16": tcgen05 path; "fp32": CUDA-core parity path; "bf16x3" /
                                                 # "bf16x6": the tcgen05 path in split-bf16 parity mode (fp32-level)
        self.cuda_graph = False                  # replay each micro-batch as a captured CUDA graph (2nd use onwards)
        # run the student on a source micro-batch and a distillation micro-batch as ONE batch (same weights, FrozenBN:
        # no cross-image coupling) -> twice the tiles per launch, half the launches; ALDI_NO_FUSE=1 is the A/B knob
        self.fuse_passes = os.environ.get("ALDI_NO_FUSE") != "1"
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("unknown StepConfig field %s" % k)
            setattr(self, k, v)

    @property
    def do_align(self):                          # aldi/trainer.py:48
        return bool(self.img_da_enabled or self.ins_da_enabled)

    def align_spec(self):
        if not self.do_align:
            return None
        return {"img": (self.img_da_input_dim, tuple(self.img_da_hidden_dims)) if self.img_da_enabled else None,
                "ins": (self.ins_da_input_dim, tuple(self.ins_da_hidden_dims)) if self.ins_da_enabled else None}

    @property
    def distill_enabled(self):                   # aldi/distill.py:140-142
        return any([self.do_hard_cls, self.do_hard_obj, self.do_hard_rpn_reg, self.do_hard_roi_reg, self.do_cls_dst,
                    self.do_obj_dst, self.do_rpn_reg_dst, self.do_roih_reg_dst])


def synthetic_state_dict_for(cfg, seed=0, convnext_layer_scale=1e-6, rel_pos_std=0.02):
    """Deterministic random-init weights of the detector a StepConfig describes, under Detectron2 key names (there is no
    network for the reference's checkpoints): heads / FPN / ResNet from arch.synthetic_state_dict plus the ConvNeXt or
    ViTDet backbone's own initialisation."""
    convnext, vitdet = cfg.backbone == "convnext", cfg.backbone in ("vitdet_b", "vitdet_l")
    sd = arch.synthetic_state_dict(seed, cfg.num_classes, align=cfg.align_spec(),
                                   bottom_up_channels=tuple(cfg.convnext_dims) if convnext else None,
                                   **(arch.VITDET_HEADS if vitdet else {}))
    if convnext:
        from .convnext import synthetic_state_dict as convnext_init
        sd.update({"backbone.bottom_up." + k: v for k, v in convnext_init(cfg.convnext_depths, cfg.convnext_dims, seed,
                                                                          convnext_layer_scale).items()})
    if vitdet:
        from . import vit
        vc = dict(vit.vit_config(cfg.backbone[-1]))
        vc.update(cfg.vit_overrides or {})
        layout = vit.ViTLayout(img_size=cfg.vit_img_size, **vc)
        sd.update({"backbone." + k: v for k, v in vit.synthetic_state_dict(layout, seed, rel_pos_std=rel_pos_std).items()})
    return sd


class MicroBatch:
    """Persistent staging buffers of one micro-batch shape: uint8 canvas, valid sizes, padded GT, the sampling seed
    and salts.  The buffers keep their addresses, so a CUDA graph captured over them can be replayed after `load`
    has refreshed their contents; per load there are n image copies plus TWO small H2D copies (int metadata, boxes)
    from pinned mirrors."""

    SITES = 3  # sampling.SITE_RPN, SITE_ROI, SITE_RPN_DISTILL

    def __init__(self, n, hp, wp, gmax, device, images=None):
        self.n, self.hp, self.wp, self.gmax = n, hp, wp, gmax
        # `images`: a (n, 3, hp, wp) view of a larger canvas shared with another micro-batch (fused student pass)
        self.images = images if images is not None else torch.zeros(n, 3, hp, wp, dtype=torch.uint8, device=device)
        ni = 1 + self.SITES * n + 2 * n + n + n * gmax
        self.h_meta = torch.zeros(ni, dtype=torch.int32).pin_memory() if device.type == "cuda" else torch.zeros(ni, dtype=torch.int32)
        self.h_boxes = torch.zeros(n, gmax, 4).pin_memory() if device.type == "cuda" else torch.zeros(n, gmax, 4)
        self.meta = torch.zeros(ni, dtype=torch.int32, device=device)
        self.boxes = torch.zeros(n, gmax, 4, device=device)
        o = 1
        self.seed = self.meta[0:1]
        self.salts = self.meta[o:o + self.SITES * n].view(self.SITES, n); o += self.SITES * n
        self.sizes = self.meta[o:o + 2 * n].view(n, 2); o += 2 * n
        counts = self.meta[o:o + n]; o += n
        classes = self.meta[o:o + n * gmax].view(n, gmax)
        self.gt = GroundTruth(self.boxes, classes, counts, gmax)
        self._full_canvas = True
        self._copied = None
        self.scratch = None
        # recorded on the compute stream after the last consumer of these buffers (initially: after the zero-fills
        # above, which a prefetch on the copy stream must not overtake)
        self.free_event = None
        if device.type == "cuda":
            self.free_event = torch.cuda.Event()
            self.free_event.record()

    @staticmethod
    def shape_key(data, with_gt):
        hs = [int(d["image"].shape[1]) if d.get("image") is not None else int(d["height"]) for d in data]
        ws = [int(d["image"].shape[2]) if d.get("image") is not None else int(d["width"]) for d in data]
        hp = (max(hs) + 31) // 32 * 32
        wp = (max(ws) + 31) // 32 * 32
        g = max([len(d["boxes"]) for d in data] + [1]) if with_gt else 1
        gmax = max(128, (g + 31) // 32 * 32)
        return len(data), hp, wp, gmax

    def load(self, data, seed, pass_id, with_gt, weak_source=None, augmenter=None):
        """Stage one micro-batch; returns the host->device bytes moved.
        Items carrying "aug_params" get their strong view DERIVED ON THE DEVICE (aldi_b200/augment.py):
        from the paired weak micro-batch `weak_source` when the item has no image of its own (unlabeled_strong), else
        from the item's own weak image staged through a scratch canvas (labeled_strong)."""
        n, gmax = self.n, self.gmax
        if self._copied is not None:
            self._copied.synchronize()   # the pinned mirrors are free again once the previous load's copies ran
        hs = [int(d["image"].shape[1]) if d.get("image") is not None else int(d["height"]) for d in data]
        ws = [int(d["image"].shape[2]) if d.get("image") is not None else int(d["width"]) for d in data]
        same = all(h == self.hp and w == self.wp for h, w in zip(hs, ws))
        if not (same and self._full_canvas):
            self.images.zero_()
        self._full_canvas = same
        nbytes = 0
        for i, d in enumerate(data):
            params = d.get("aug_params")
            img = d.get("image")
            if params is None:
                self.images[i, :, :hs[i], :ws[i]].copy_(img, non_blocking=True)
            else:
                assert augmenter is not None, "items with aug_params need a StrongAugmenter"
                if img is not None:
                    if self.scratch is None:
                        self.scratch = torch.zeros_like(self.images[0])
                    self.scratch[:, :hs[i], :ws[i]].copy_(img, non_blocking=True)
                    src = self.scratch
                else:
                    assert weak_source is not None, "an image-less strong item needs its weak micro-batch"
                    src = weak_source.images[i]
                    if weak_source._copied is not None:
                        # the weak images may have been staged on ANOTHER stream (this load can be a prefetch on the copy
                        # stream): derive the strong view only after their copies have landed
                        torch.cuda.current_stream().wait_event(weak_source._copied)
                augmenter.apply(src, self.images[i], params, valid_hw=(hs[i], ws[i]))
            if img is not None and not img.is_cuda:
                nbytes += img.numel()
        m = self.h_meta
        m.zero_()
        m[0] = seed - (1 << 32) if seed >= (1 << 31) else seed
        o = 1
        for site in range(self.SITES):
            for i in range(n):
                v = sampling.make_salt(pass_id, site, i)
                m[o + site * n + i] = v - (1 << 32) if v >= (1 << 31) else v
        o += self.SITES * n
        for i in range(n):
            m[o + 2 * i], m[o + 2 * i + 1] = hs[i], ws[i]
        o += 2 * n
        if with_gt:
            self.h_boxes.zero_()
            for i, d in enumerate(data):
                k = len(d["boxes"])
                assert k <= gmax
                m[o + i] = k
                if k:
                    self.h_boxes[i, :k] = d["boxes"].float()
                    m[o + n + i * gmax:o + n + i * gmax + k] = d["classes"].to(torch.int32)
            self.boxes.copy_(self.h_boxes, non_blocking=True)
            nbytes += self.h_boxes.numel() * 4
        self.meta.copy_(m, non_blocking=True)
        nbytes += m.numel() * 4
        if self.images.is_cuda:
            self._copied = torch.cuda.Event()
            self._copied.record()
        return nbytes


class GroundTruth:
    def __init__(self, boxes, classes, counts, gmax, scores=None):
        self.boxes, self.classes, self.counts, self.gmax, self.scores = boxes, classes, counts, gmax, scores

    @staticmethod
    def from_host(boxes_list, classes_list, device):
        n = len(boxes_list)
        gmax = max(32, (max([len(b) for b in boxes_list] + [1]) + 31) // 32 * 32)
        b = torch.zeros(n, gmax, 4)
        c = torch.zeros(n, gmax, dtype=torch.int32)
        cnt = torch.zeros(n, dtype=torch.int32)
        for i, (bb, cc) in enumerate(zip(boxes_list, classes_list)):
            k = len(bb)
            if k:
                b[i, :k] = bb.float()
                c[i, :k] = cc.to(torch.int32)
            cnt[i] = k
        return GroundTruth(b.to(device, non_blocking=True), c.to(device, non_blocking=True),
                           cnt.to(device, non_blocking=True), gmax)


class B200TrainStep:
    def __init__(self, cfg, state_dict, device="cuda:0", teacher_state_dict=None, process_group=None):
        self.cfg = cfg
        self.device = torch.device(device)
        if cfg.dtype not in ("bf16", "fp32", "bf16x3", "bf16x6"):
            raise ValueError("StepConfig.dtype %r: bf16 | fp32 | bf16x3 | bf16x6" % (cfg.dtype,))
        self.dtype = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
        self.dtc = _l.BF16 if cfg.dtype == "bf16" else _l.F32
        # "bf16x3" / "bf16x6": fp32 activations, every GEMM on the tcgen05 kernels as 3 / 6 bf16 product terms
        split_parts = {"bf16x3": 2, "bf16x6": 3}.get(cfg.dtype, 0)
        _l.load()  # fail loudly if the CUDA library is missing
        convnext = cfg.backbone == "convnext"
        vitdet = cfg.backbone in ("vitdet_b", "vitdet_l")
        if cfg.backbone not in ("resnet50", "convnext", "vitdet_b", "vitdet_l"):
            raise NotImplementedError("MODEL.BACKBONE %s: ResNet-50-FPN, ConvNeXt-FPN and ViTDet-B/L are built" % cfg.backbone)
        if vitdet:
            if cfg.do_align:
                raise NotImplementedError("domain-alignment discriminators on the ViTDet heads are not built")
            if cfg.dtype not in ("bf16", "fp32"):
                raise NotImplementedError("ViTDet runs in bf16 (tcgen05) or fp32 (parity mode)")
        self.layout = FlatLayout(cfg.num_classes, align=cfg.align_spec(),
                                 bottom_up_channels=tuple(cfg.convnext_dims) if convnext else None,
                                 head=arch.VITDET_HEADS if vitdet else None)
        self.det = Detector(cfg.num_classes, cfg.anchor_sizes, cfg.pixel_mean, cfg.pixel_std)
        flat = self.layout.pack_state_dict(state_dict).to(self.device)
        tsd = state_dict if teacher_state_dict is None else teacher_state_dict
        tflat = flat.clone() if teacher_state_dict is None else self.layout.pack_state_dict(tsd).to(self.device)
        bu_s = bu_t = None
        if convnext:
            from .convnext import ConvNeXtBackbone
            pre = "backbone.bottom_up."

            def bottom_up(sd):
                return ConvNeXtBackbone({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, depths=cfg.convnext_depths,
                                        dims=cfg.convnext_dims, drop_path_rate=cfg.convnext_drop_path, dtype=cfg.dtype,
                                        device=self.device, pixel_mean=cfg.pixel_mean, pixel_std=cfg.pixel_std)

            bu_s, bu_t = bottom_up(state_dict), bottom_up(tsd)
        if vitdet:
            from .vit import ViTDetBackbone
            pre = "backbone."

            def pyramid(sd):
                return ViTDetBackbone({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, size=cfg.backbone[-1],
                                      dtype=cfg.dtype, device=self.device, pixel_mean=cfg.pixel_mean, pixel_std=cfg.pixel_std,
                                      img_size=cfg.vit_img_size, **(cfg.vit_overrides or {}))

            bu_s, bu_t = pyramid(state_dict), pyramid(tsd)
        # DropPath masks (host-drawn, aldi/backbone.py:176-181): every data-parallel rank draws its own
        self.keep_rng = torch.Generator().manual_seed(dist.get_rank(process_group) if process_group is not None else 0)
        self.keep_override = None                               # test seam: list of per-forward mask lists
        self.student = DetectorWeights(self.layout, flat, self.dtype, bottom_up=bu_s, split_parts=split_parts)
        self.teacher = DetectorWeights(self.layout, tflat, self.dtype, bottom_up=bu_t, split_parts=split_parts)
        self.student.enable_dgrad()
        self.nt = self.layout.num_trainable
        self.grad = torch.zeros(self.nt, device=self.device)
        self.momentum_buf = torch.zeros(self.nt, device=self.device)
        if cfg.optimizer.upper() == "ADAMW":
            self.exp_avg_sq = torch.zeros(self.nt, device=self.device)
            if bu_s is not None:
                self.bu_exp_avg, self.bu_exp_avg_sq = torch.zeros_like(bu_s.flat), torch.zeros_like(bu_s.flat)
        elif cfg.optimizer.upper() != "SGD":
            raise ValueError("Unsupported optimizer/backbone combination {} {}.".format(cfg.optimizer, cfg.backbone))
        elif bu_s is not None:
            self.bu_momentum = torch.zeros_like(bu_s.flat)
        self.student.refresh()
        self.teacher.refresh()
        self.loss_acc = torch.zeros(LOSS_SLOTS, device=self.device)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.pg = process_group
        self.reducer = GradReducer(self.layout, self.grad, process_group)
        if self.reducer.active:
            # under data parallelism NCCL kernels share the SMs with the step: early-launched dependents only park there
            # (measured at 8 GPUs: 2103 images/s without programmatic dependent launch, 2095 with)
            _l.load().aldi_set_pdl(0)
        self._last_backward = False
        self._mb_cache = {}        # (kind, n, hp, wp, gmax) -> MicroBatch staging buffers
        self._graphs = {}          # key -> None (seen once, ran eagerly) | torch.cuda.CUDAGraph
        self._graph_pool = None
        self._capture_hook = None
        self.profile_spin_cycles = 0
        self._h2d_stream = None
        self.graph_replays = 0
        self.iter = 0
        self.h2d_bytes = 0
        self.last_pseudo = None
        self.debug = {}
        # test seam (cf. the reference's DEBUG dict, aldi/trainer.py:24-25): pseudo labels to use INSTEAD of the
        # device-computed ones, one GroundTruth per distillation micro-batch
        self.pseudo_override = None
        self.pseudo_log = []
        # test seam of the same kind for the student's RPN proposals: {pass_id: [(k_i, 4) boxes per image]} to use INSTEAD
        # of the device-computed ones (which are kept in `proposal_log` for comparison).  Selections are discontinuous: at
        # 1024x2048 the top-2000 cut falls among ~400k scores a few 1e-6 apart, so two correct implementations differ in
        # a handful of boxes; everything downstream is then compared on identical inputs.
        self.proposal_override = None
        self.proposal_log = {}
        # {"labeled": StrongAugmenter, "unlabeled": StrongAugmenter}: set to derive strong views on the device from
        # items that carry "aug_params" 
        self.augmenters = {}
        self._teacher_out = None   # outputs of the last teacher pass, consumed by the fused student pass
        self._teacher_outs = {}    # ... per captured teacher graph: a replay writes into THAT capture's tensors
        self._mask_slots, self._mask_events = {}, {}     # DropPath mask blocks per micro-batch body (see _keep_masks)
        self._mask_key, self._mask_idx = None, 0

    # ---- aldi/ema.py:52-57 -------------------------------------------------------------------------
    def ema_update(self, it):
        alpha = 0.0 if it <= self.cfg.ema_start_iter else self.cfg.ema_alpha
        ops.ema_update(self.teacher.flat, self.student.flat, alpha)
        self.teacher.refresh()
        if self.teacher.bottom_up is not None:
            ops.ema_update(self.teacher.bottom_up.flat, self.student.bottom_up.flat, alpha)
            self.teacher.bottom_up.refresh()

    # ---- aldi/trainer.py:28-117 ----------------------------------------------------------------------
    def run_model(self, data):
        labeled_weak, labeled_strong, unlabeled_weak, unlabeled_strong = data
        cfg = self.cfg
        mb = cfg.ims_per_gpu
        do_weak, do_strong = labeled_weak is not None, labeled_strong is not None
        do_distill = cfg.distill_enabled and unlabeled_weak is not None
        total = sum(len(s or []) for s in (labeled_weak, labeled_strong, unlabeled_weak))
        accum = total // mb
        assert accum >= 1, "SOLVER.IMS_PER_GPU larger than the per-GPU batch"
        gscale = 1.0 / accum
        self.loss_acc.zero_()
        self.h2d_bytes = 0
        self.seed = random.randint(0, 2 ** 32 - 1)          # aldi/helpers.py:19-23
        self.seed_log = {}
        self.pseudo_log = []
        out_keys = []
        # ---- plan the micro-batches (same order and seed draws as the reference loops), then run them with the
        # NEXT micro-batch's host->device staging overlapped on a copy stream whenever it uses other buffers
        plan = []
        do_align = cfg.do_align
        da_on = (cfg.img_da_enabled, cfg.ins_da_enabled)
        for tag, d in (("source_weak", labeled_weak), ("source_strong", labeled_strong)):
            if d is None:
                continue
            base = SLOT_BASE[tag]
            for i in range(0, len(d), mb):
                plan.append({"kind": "source", "parts": [("source", d[i:i + mb], True)], "seed": self.seed, "base": base})
            out_keys += [(k + "_" + tag, base + j) for j, k in enumerate(SRC_KEYS)]
            out_keys += [(k + "_" + tag, base + 4 + j) for j, (k, on) in enumerate(zip(DA_KEYS, da_on)) if on]
        if do_align:
            # weakly-augmented target imagery, domain label 0, alignment losses only (aldi/trainer.py:107-109)
            assert unlabeled_weak is not None, "domain alignment needs unlabeled (target) data"
            base = SLOT_BASE["target_weak"]
            for i in range(0, len(unlabeled_weak), mb):
                plan.append({"kind": "align_target", "parts": [("target", unlabeled_weak[i:i + mb], False)],
                             "seed": self.seed, "base": base})
            out_keys += [(k + "_target_weak", base + j) for j, (k, on) in enumerate(zip(DA_KEYS, da_on)) if on]
        if do_distill:
            assert len(unlabeled_weak) == len(unlabeled_strong), "Teacher and student data must be the same length."
            base = SLOT_BASE["distill"]
            for i in range(0, len(unlabeled_weak), mb):
                seed = random.randint(0, 2 ** 32 - 1)       # seeder.reset_seed(), aldi/distill.py:150
                plan.append({"kind": "distill", "seed": seed, "base": base,
                             "parts": [("weak", unlabeled_weak[i:i + mb], False), ("strong", unlabeled_strong[i:i + mb], False)]})
            out_keys += [(k + "_distill", base + j) for j, k in enumerate(SRC_KEYS)]
            soft_on = (cfg.do_obj_dst, cfg.do_rpn_reg_dst, cfg.do_cls_dst, cfg.do_roih_reg_dst)
            out_keys += [(k + "_distill", base + 4 + j) for j, (k, on) in enumerate(zip(SOFT_KEYS, soft_on)) if on]
            if do_align:
                # the student's forward inside the distiller returns the `_da` placeholder (aldi/align.py:91-100),
                # which passes the distill key filter `k != "_"` (aldi/trainer.py:113) as a zero
                out_keys.append(("_da_distill", SLOT_ZERO))
        for i, item in enumerate(plan):
            item["pass_id"] = 100 + i if item["kind"] == "distill" else i
            item["keys"] = [(kind,) + MicroBatch.shape_key(d, gt) for kind, d, gt in item["parts"]]
            self.seed_log[item["pass_id"]] = item["seed"]
        if cfg.fuse_passes and do_distill and not do_align and self.student.bottom_up is None:
            plan = self._fuse_plan(plan)
        last_bwd = max(i for i, it in enumerate(plan) if it["kind"] != "teacher")
        bodies = {"source": self._source_body, "distill": self._distill_body, "align_target": self._align_target_body,
                  "teacher": self._teacher_body, "fused": self._fused_body}
        for i, item in enumerate(plan):
            self.seed = item["seed"]
            self._last_backward = i == last_bwd
            mbs = item.get("staged") or self._stage(item, None)
            if item["kind"] == "teacher":
                item["fused"]["weak_mb"] = mbs[0]      # the strong views may be derived from it on the device
            body = bodies[item["kind"]]
            self._run((item["kind"],) + tuple(item["keys"]) + (gscale, self._last_backward, item["base"]),
                      lambda: body(*mbs, gscale, item["pass_id"], item["base"]))
            if self.device.type == "cuda":
                for m in mbs:
                    m.free_event = torch.cuda.Event()
                    m.free_event.record()
                nxt = plan[i + 1] if i + 1 < len(plan) else None
                if nxt is not None and not (set(nxt["keys"]) & set(item["keys"])):
                    nxt["staged"] = self._stage(nxt, self._copy_stream())
        self._out_keys = out_keys
        return LossDict(self, out_keys)

    def _fuse_plan(self, plan):
        """Pair the k-th source micro-batch with the k-th distillation micro-batch: [teacher(weak_k), fused(source_k +
        strong_k)].  Pass ids, seeds and salts stay those of the reference-order plan, so every sample drawn is the
        same; only the order in which the gradient contributions are summed changes.  Pairs whose canvases differ
        in size, and unpaired micro-batches, run through the unfused bodies."""
        src = [it for it in plan if it["kind"] == "source"]
        dst = [it for it in plan if it["kind"] == "distill"]
        out, used = [], set()
        for s_it, d_it in zip(src, dst):
            (_, sdata, _), = s_it["parts"]
            (_, wdata, _), (_, tdata, _) = d_it["parts"]
            ks, kt = MicroBatch.shape_key(sdata, True), MicroBatch.shape_key(tdata, False)
            if ks[1:3] != kt[1:3]:
                continue
            used.update((id(s_it), id(d_it)))
            canvas = ("canvas", ks[0], kt[0], ks[1], ks[2])
            fused = {"kind": "fused", "seed": d_it["seed"], "base": (s_it["base"], d_it["base"]),
                     "pass_id": (s_it["pass_id"], d_it["pass_id"]), "canvas": canvas,
                     "parts": [("fsource", sdata, True, s_it["seed"], s_it["pass_id"], 0),
                               ("fstrong", tdata, False, d_it["seed"], d_it["pass_id"], ks[0])],
                     "keys": [("fsource",) + ks + (kt[0],), ("fstrong",) + kt + (ks[0],)]}
            out.append({"kind": "teacher", "seed": d_it["seed"], "base": d_it["base"], "pass_id": d_it["pass_id"],
                        "parts": [("weak", wdata, False)], "keys": [d_it["keys"][0]], "fused": fused})
            out.append(fused)
        return out + [it for it in plan if id(it) not in used]

    def losses_to_host(self, out_keys=None):
        vals = self.loss_acc.cpu()  # one D2H read of the step's loss vector (the reference syncs per loss)
        flag = int(self.err_flag.item())
        if flag & 2:
            raise _l.AldiError("sparse RPN backward: more non-zero head-gradient rows than its capacity (4 x RPN.BATCH_SIZE_PER_IMAGE "
                               "per image); run with ALDI_SPARSE_RPN_BWD=0")
        if flag:
            raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
        return {k: float(vals[j]) for k, j in (out_keys or self._out_keys)}

    # ---- one source micro-batch: student forward + backward with hard losses --------------------------
    def _copy_stream(self):
        if self._h2d_stream is None:
            self._h2d_stream = torch.cuda.Stream(device=self.device)
        return self._h2d_stream

    def _stage(self, item, stream):
        """Stage the inputs of one planned micro-batch into their MicroBatch buffers; `stream` = the copy stream for a
        prefetch (ordered after the buffers' last consumer, the compute stream then waits for the copies) or None
        for plain in-order staging on the compute stream."""
        out = []
        for key, part in zip(item["keys"], item["parts"]):
            kind, data, with_gt = part[:3]
            seed, pass_id = (part[3], part[4]) if len(part) > 3 else (item["seed"], item["pass_id"])
            mb = self._mb_cache.get(key)
            if mb is None:
                view = None
                if len(part) > 5:
                    # fused student pass: both micro-batches are views of ONE image canvas, source images first
                    ck = item["canvas"]
                    canvas = self._mb_cache.get(ck)
                    if canvas is None:
                        canvas = self._mb_cache[ck] = torch.zeros(ck[1] + ck[2], 3, ck[3], ck[4], dtype=torch.uint8,
                                                                  device=self.device)
                    view = canvas[part[5]:part[5] + key[1]]
                mb = self._mb_cache[key] = MicroBatch(*key[1:5], self.device, images=view)
            aug = self.augmenters.get("labeled" if with_gt else "unlabeled")
            weak = out[0] if (kind == "strong" and out) else item.get("weak_mb") if kind == "fstrong" else None
            if stream is None:
                self.h2d_bytes += mb.load(data, seed, pass_id, with_gt, weak, aug)
            else:
                with torch.cuda.stream(stream):
                    if mb.free_event is not None:
                        stream.wait_event(mb.free_event)
                    self.h2d_bytes += mb.load(data, seed, pass_id, with_gt, weak, aug)
                    ready = torch.cuda.Event()
                    ready.record(stream)
                torch.cuda.current_stream().wait_event(ready)
            out.append(mb)
        return out

    def _run(self, key, fn):
        """Run one micro-batch body: eagerly the first time a (shape, schedule) key is seen, from then on as
        captured CUDA graphs (cfg.cuda_graph).  All inputs live in MicroBatch buffers with fixed addresses, the
        sampling seed and salts included, so a replay sees the new step's data.  The step's LAST backward under
        data parallelism is captured as a CHAIN of graphs cut where a gradient bucket becomes final: the NCCL
        all-reduce of that bucket is issued eagerly between two replays and overlaps the next segment."""
        self._mask_key, self._mask_idx = key, 0
        if not (self.cfg.cuda_graph and self.device.type == "cuda") or self.debug is not None or self.pseudo_override \
                or self.proposal_override or self.keep_override is not None:
            if self.profile_spin_cycles:
                # profiling aid: park the GPU on a spin kernel while the host queues this micro-batch, so the kernels
                # then run back to back (warm L2, no launch gaps) and per-launch CUDA events time exactly their durations
                torch.cuda._sleep(int(self.profile_spin_cycles))
            return fn()
        chain = self._graphs.get(key, False)
        if chain is False:
            self._graphs[key] = None
            return fn()
        if chain is None:
            chain = self._graphs[key] = self._capture(fn)
            if key[0] == "teacher":
                self._teacher_outs[key] = self._teacher_out
        if key[0] == "teacher":
            # the fused student pass that follows reads `_teacher_out`: point it at the tensors THIS graph writes (another
            # shape's teacher pass may have run, eagerly or captured, in between)
            self._teacher_out = self._teacher_outs[key]
        self._redraw_masks(key)
        for seg in chain:
            if isinstance(seg, str):
                self.reducer.ready(seg)
            else:
                seg.replay()
                self.graph_replays += 1
        if key in self._mask_slots:
            self._mask_events[key] = torch.cuda.Event()
            self._mask_events[key].record()

    def _capture(self, fn):
        import gc
        import warnings
        warnings.filterwarnings("ignore", message="The CUDA Graph is empty")  # the tail after the last bucket cut
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
            self._capture_stream = torch.cuda.Stream(device=self.device)
        chain = []
        # ALDI_GRAPH_SPLIT=1: cut the chain even on one GPU (test knob for the segmented capture)
        split = self._last_backward and (self.reducer.active or os.environ.get("ALDI_GRAPH_SPLIT") == "1")
        torch.cuda.synchronize()
        gc.collect()
        cs = self._capture_stream
        cs.wait_stream(torch.cuda.current_stream())
        # thread-local capture mode: under torch.distributed the NCCL watchdog thread polls its work events with
        # cudaEventQuery, which the default ("global") mode forbids in EVERY thread while a capture is open -- the
        # watchdog then dies with "operation not permitted when stream is capturing" and takes the process down
        # (SIGABRT); the calls of the capturing thread itself are still checked
        mode = os.environ.get("ALDI_CAPTURE_MODE", "thread_local")
        with torch.cuda.stream(cs):
            seg = torch.cuda.CUDAGraph()
            seg.capture_begin(pool=self._graph_pool, capture_error_mode=mode)
            state = {"seg": seg}

            def cut(tag):
                state["seg"].capture_end()
                chain.extend([state["seg"], tag])
                state["seg"] = torch.cuda.CUDAGraph()
                state["seg"].capture_begin(pool=self._graph_pool, capture_error_mode=mode)

            self._capture_hook = cut if split else None
            try:
                fn()
            finally:
                self._capture_hook = None
                state["seg"].capture_end()
            chain.append(state["seg"])
        torch.cuda.current_stream().wait_stream(cs)
        return chain

    def _source_body(self, b, gscale, pass_id, base=SLOT_BASE["source_strong"]):
        fw = self._student_forward(b, b.gt, pass_id, want_rpn_labels=True)
        self._source_finish(fw, b, {k: gscale for k in SRC_KEYS + ("da",)}, self.loss_acc, base, labeled=True)

    def _source_finish(self, fw, b, w, loss_vec, base, labeled=True, do_align=True):
        """Hard losses of a finished student forward and the explicit backward.  w: weight per loss key (`loss_cls`,
        `loss_box_reg`, `loss_rpn_cls`, `loss_rpn_loc`, `da`): value written = weight * loss, gradient scaled alike
        (the step passes 1/num_grad_accum for all of them; the module facade passes what autograd hands back)."""
        d_rpn, dpred, align = self._source_losses(fw, b, w, loss_vec, base, labeled, do_align, self.grad)
        self.det.backward(self.student, self.grad, fw["feats"], fw["saved"], fw["rpn_ts"], d_rpn, fw["lv"], fw["head_saved"],
                          dpred, fw["rois"], fw["roi_batch"], on_ready=self._bucket_ready(), **self._bwd_kw(), align=align)

    def _source_losses(self, fw, b, w, loss_vec, base, labeled, do_align, G):
        """-> (d_rpn, dpred, align context); G: where the discriminators' last-layer gradients go (they are produced by
        the same kernel as the loss)."""
        cfg, det = self.cfg, self.det
        n = b.n
        if "labels" not in fw:        # target_weak alignment pass: no detection losses (aldi/trainer.py:107-109)
            return None, None, (self._align(fw, labeled, w["da"], loss_vec[base:base + 2], G) if do_align else None)
        d_rpn = torch.zeros(n, fw["lv"].total_locs, 64, device=self.device, dtype=self.dtype)
        ops.call("aldi_rpn_loss", fw["rpn_out"], _l.ctypes.byref(fw["lv"]), n, fw["labels"], fw["matched"], b.gt.boxes,
                 b.gt.counts, b.gt.gmax, cfg.rpn_batch, w["loss_rpn_cls"], w["loss_rpn_loc"], 1.0, d_rpn, self.dtc, 64, 0,
                 loss_vec[base + 2:base + 4])
        m = n * cfg.roi_batch
        dpred = torch.zeros(m, 64, device=self.device, dtype=self.dtype)
        ops.call("aldi_roi_loss", fw["pred"], det.PRED_CH, m, cfg.num_classes, fw["roi_class"], fw["rois"], fw["roi_gt"],
                 fw["roi_count"], n, ops.host_floats((10.0, 10.0, 5.0, 5.0)), w["loss_cls"], w["loss_box_reg"], 1.0, dpred,
                 self.dtc, 64, loss_vec[base:base + 2])
        align = self._align(fw, labeled, w["da"], loss_vec[base + 4:base + 6], G) if do_align else None
        return d_rpn, dpred, align

    def _align(self, fw, labeled, gscale, loss_out, G=None):
        """AlignMixin.forward(do_align=True) on a finished student forward (aldi/align.py:74-90)."""
        if not self.cfg.do_align:
            return None
        saved = self.det.align_forward(self.student, self.grad if G is None else G, fw["feats"], fw["head_saved"][2],
                                       fw["roi_count"], self.cfg.roi_batch, self.cfg, labeled, gscale, loss_out)
        saved["img_layer"] = self.cfg.img_da_layer
        return saved

    def _align_target_body(self, b, gscale, pass_id, base=SLOT_BASE["target_weak"]):
        """aldi/trainer.py:107-109: model(unlabeled_weak, labeled=False, do_align=True) — a training-mode student
        forward on target images with EMPTY ground truth (aldi/dataloader.py:21-30), of which only the `_da_` losses
        are kept and back-propagated (the detection losses are multiplied by 0, aldi/trainer.py:75-77)."""
        fw = self._student_forward(b, b.gt, pass_id, want_rpn_labels=False)
        self._source_finish(fw, b, {"da": gscale}, self.loss_acc, base, labeled=False)

    def _student_forward(self, b, gt, pass_id, want_rpn_labels):
        cfg, det, W = self.cfg, self.det, self.student
        n = b.n
        feats, saved = det.backbone(W, b.images, b.sizes, save=True, keep_masks=self._keep_masks(W, n))
        lv = det.levels(feats)
        rpn_out, rpn_ts = det.rpn_head(W, feats, lv, save=True)
        out = {"feats": feats, "saved": saved, "lv": lv, "rpn_out": rpn_out, "rpn_ts": rpn_ts}
        if want_rpn_labels:
            out["labels"], out["matched"], out["rpn_stats"] = self._label_anchors(lv, b, gt, sampling.SITE_RPN)
        props = det.proposals(rpn_out, lv, b.sizes, cfg.rpn_pre_topk[0], cfg.rpn_post_topk[0], cfg.rpn_nms_thresh,
                              self.err_flag)
        props = self._override_proposals(props, [(pass_id, n)])
        out["props"] = props
        m = n * cfg.roi_batch
        rois = torch.empty(m, 4, device=self.device)
        roi_gt = torch.empty(m, 4, device=self.device)
        roi_batch = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_class = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_src = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_count = torch.zeros(n, dtype=torch.int32, device=self.device)
        roi_stats = torch.zeros(n, 2, dtype=torch.int32, device=self.device)
        ops.call("aldi_roi_label_sample", props["boxes"], props["count"], props["boxes"].shape[1], n, gt.boxes,
                 gt.classes, gt.counts, gt.gmax, cfg.roi_iou, cfg.num_classes, cfg.roi_batch, cfg.roi_pos_fraction,
                 b.seed, b.salts[sampling.SITE_ROI], 1, rois, roi_batch, roi_class, roi_gt, roi_src, roi_count, roi_stats)
        pred, head_saved = det.box_head(W, feats, rois, roi_batch, save=True)
        out.update(rois=rois, roi_gt=roi_gt, roi_batch=roi_batch, roi_class=roi_class, roi_src=roi_src,
                   roi_count=roi_count, roi_stats=roi_stats, pred=pred, head_saved=head_saved)
        return out

    def _override_proposals(self, props, passes):
        """Test seam (see `proposal_override`): log the device's proposals per pass, substitute the given ones.
        passes: [(pass_id, images)] in row order of `props`."""
        if self.proposal_override is None:
            return props
        boxes, count = props["boxes"].clone(), props["count"].clone()
        row = 0
        for pid, k in passes:
            self.proposal_log[pid] = (props["boxes"][row:row + k].clone(), props["count"][row:row + k].clone())
            for j in range(k):
                b = self.proposal_override[pid][j].to(self.device, torch.float32)
                boxes[row + j].zero_()
                boxes[row + j, :b.shape[0]] = b
                count[row + j] = b.shape[0]
            row += k
        out = dict(props)
        out["boxes"], out["count"] = boxes, count
        return out

    def _keep_masks(self, W, n):
        """DropPath factors of one training-mode forward of a ConvNeXt bottom-up (None for the ResNet).
        Drawn on the host like `torch.bernoulli_` in the reference (aldi/backbone.py:176-181) but DEVICE-RESIDENT: every
        mask request of a micro-batch body owns a pinned host block and a device block; the body copies one onto the
        other (a memcpy node when the body is captured), and `_redraw_masks` refreshes the pinned blocks before each graph
        replay -- so a replayed ConvNeXt step sees fresh masks and nothing about DropPath forces eager execution."""
        if W.bottom_up is None:
            return None
        if self.keep_override is not None:
            return self.keep_override.pop(0)
        rates = W.bottom_up.drop_rates
        slots = self._mask_slots.setdefault(self._mask_key, [])
        i = self._mask_idx
        self._mask_idx += 1
        if i == len(slots):
            host = torch.zeros(len(rates), n)
            slots.append((host.pin_memory() if self.device.type == "cuda" else host,
                          torch.zeros(len(rates), n, device=self.device), rates))
        host, dev, _ = slots[i]
        self._draw_into(host, rates)
        dev.copy_(host, non_blocking=True)
        return [None if r <= 0 else dev[b] for b, r in enumerate(rates)]

    def _draw_into(self, host, rates):
        n = host.shape[1]
        for b, r in enumerate(rates):
            if r > 0:
                host[b] = (torch.rand(n, generator=self.keep_rng) < (1 - r)).float() / (1 - r)

    def _redraw_masks(self, key):
        """Before a graph replay: new DropPath draws into the pinned blocks the captured copies read."""
        slots = self._mask_slots.get(key)
        if not slots:
            return
        ev = self._mask_events.get(key)
        if ev is not None:
            ev.synchronize()          # the previous replay's copies have read the pinned blocks
        for host, _, rates in slots:
            self._draw_into(host, rates)

    def _label_anchors(self, lv, b, gt, site):
        cfg, n = self.cfg, b.n
        total = lv.total_locs * lv.num_anchors
        labels = torch.empty(n, total, dtype=torch.int8, device=self.device)
        matched = torch.empty(n, total, dtype=torch.int32, device=self.device)
        stats = torch.zeros(n, 2, dtype=torch.int32, device=self.device)
        wsb = int(_l.load().aldi_rpn_label_workspace_bytes(n, gt.gmax))
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        ops.call("aldi_rpn_label_anchors", _l.ctypes.byref(lv), n, gt.boxes, gt.counts, gt.gmax, cfg.rpn_iou[0],
                 cfg.rpn_iou[1], cfg.rpn_batch, cfg.rpn_pos_fraction, b.seed, b.salts[site], ws, wsb, labels, matched, stats)
        return labels, matched, stats

    # ---- teacher: trunk + RPN once, eval-mode detections -> pseudo labels --------------------------------
    def teacher_forward(self, b, train_mode=False):
        """train_mode: the teacher's second forward inside the distiller runs in TRAINING mode (aldi/distill.py:153-162),
        which matters only for a bottom-up with DropPath; the FrozenBN ResNet gives identical features in both modes."""
        cfg, det, W = self.cfg, self.det, self.teacher
        feats, _ = det.backbone(W, b.images, b.sizes, save=False, keep_masks=self._keep_masks(W, b.n) if train_mode else None)
        lv = det.levels(feats)
        rpn_out, _ = det.rpn_head(W, feats, lv, save=False)
        return feats, lv, rpn_out

    def pseudo_label(self, b, feats, lv, rpn_out, score_thresh=None, W=None):
        """aldi/pseudolabeler.py:15-30: teacher.inference(do_postprocess=False) then scores > threshold."""
        cfg, det = self.cfg, self.det
        W = W if W is not None else self.teacher
        n = b.n
        props = det.proposals(rpn_out, lv, b.sizes, cfg.rpn_pre_topk[1], cfg.rpn_post_topk[1], cfg.rpn_nms_thresh)
        p = props["boxes"].shape[1]
        roi_batch = torch.arange(n, dtype=torch.int32, device=self.device).repeat_interleave(p)
        pred, _ = det.box_head(W, feats, props["boxes"].view(n * p, 4), roi_batch, save=False)
        # greedy NMS in descending score order: a detection's fate depends only on higher-scoring ones, so
        # candidates below the pseudo-label threshold can never change the thresholded result
        thr = max(cfg.test_score_thresh, cfg.pseudo_threshold) if score_thresh is None else score_thresh
        dets = det.detections(pred, props, b.sizes, thr, cfg.test_nms_thresh, cfg.test_topk)
        gmax = 128
        gb = torch.zeros(n, gmax, 4, device=self.device)
        gc = torch.zeros(n, gmax, dtype=torch.int32, device=self.device)
        gs = torch.zeros(n, gmax, device=self.device)
        cnt = torch.zeros(n, dtype=torch.int32, device=self.device)
        ops.call("aldi_pseudo_label_threshold", dets["boxes"], dets["scores"], dets["cats"], dets["count"],
                 cfg.test_topk, n, cfg.pseudo_threshold, gb, gc, gs, cnt, gmax)
        return GroundTruth(gb, gc, cnt, gmax, gs), dets

    # ---- model.inference(batched_inputs, do_postprocess) (detectron2 GeneralizedRCNN.inference) -----------------
    def inference(self, batched_inputs, which="teacher", do_postprocess=True):
        """Eval-mode detections of the EMA teacher (what `EMA.inference` / the periodic evaluation of the teacher run,
        aldi/ema.py:59-60, aldi/trainer.py:173-185) or of the student: RPN test top-k, box head, score threshold
        TEST score 0.05, per-class NMS 0.5, top 100.  Returns detectron2's output format: a list of
        {"instances": Instances(pred_boxes, scores, pred_classes)} rescaled to each input's (height, width)
        (detector_postprocess) when do_postprocess, else the bare Instances in the network input frame
        (aldi/pseudolabeler.py:21 calls it that way).  ONE device->host copy per micro-batch."""
        from .structures import Boxes, Instances
        cfg, det = self.cfg, self.det
        W = self.teacher if which == "teacher" else self.student
        out = []
        for i in range(0, len(batched_inputs), cfg.ims_per_gpu):
            data = batched_inputs[i:i + cfg.ims_per_gpu]
            key = ("infer",) + MicroBatch.shape_key(data, False)
            mb = self._mb_cache.get(key)
            if mb is None:
                mb = self._mb_cache[key] = MicroBatch(*key[1:], self.device)
            mb.load(data, 0, 0, False)
            feats, _ = det.backbone(W, mb.images, mb.sizes, save=False)
            lv = det.levels(feats)
            rpn_out, _ = det.rpn_head(W, feats, lv, save=False)
            _, dets = self.pseudo_label(mb, feats, lv, rpn_out, score_thresh=cfg.test_score_thresh, W=W)
            cnt = dets["count"].cpu().tolist()
            boxes, scores, cats = dets["boxes"].cpu(), dets["scores"].cpu(), dets["cats"].cpu()
            for j, d in enumerate(data):
                k = cnt[j]
                h_in, w_in = int(mb.h_meta[1 + MicroBatch.SITES * mb.n + 2 * j]), int(mb.h_meta[2 + MicroBatch.SITES * mb.n + 2 * j])
                inst = Instances((h_in, w_in), pred_boxes=Boxes(boxes[j, :k].clone()), scores=scores[j, :k].clone(),
                                 pred_classes=cats[j, :k].long())
                if not do_postprocess:
                    out.append(inst)
                    continue
                oh, ow = int(d.get("height", h_in)), int(d.get("width", w_in))
                out.append({"instances": detector_postprocess(inst, oh, ow)})
        return out

    # ---- fused schedule: teacher pass, then ONE student pass over (source + strong target) images --------------------
    def _teacher_body(self, bw, gscale, pass_id, base):
        """First half of aldi/distill.py:144-168 for one micro-batch: teacher trunk + RPN once, pseudo labels.  The
        outputs stay referenced (`_teacher_out`) because the student pass that consumes them is a separate graph."""
        t_feats, t_lv, t_rpn_out = self.teacher_forward(bw)
        pseudo, _ = self.pseudo_label(bw, t_feats, t_lv, t_rpn_out)
        self.last_pseudo = pseudo
        if self.debug is not None:
            self.pseudo_log.append(pseudo)
        if self.pseudo_override is not None:
            pseudo = self.pseudo_override[len(self.pseudo_log) - 1]
        self._teacher_out = {"feats": t_feats, "lv": t_lv, "rpn_out": t_rpn_out, "pseudo": pseudo}

    def _fused_body(self, bsrc, bstr, gscale, pass_ids, bases):
        """Student forward + backward of a source micro-batch (hard losses against GT) and a distillation micro-batch
        (soft losses against the teacher) as ONE batch of ns + nt images: the trunk, FPN, RPN head and box head see
        twice the tiles per launch; labelling, sampling and the losses run per half with that half's seed, salts,
        ground truth and normalisers, so every loss and gradient equals the two separate passes of the reference
        (aldi/trainer.py:86-97) up to the order of the floating-point sums."""
        cfg, det, W = self.cfg, self.det, self.student
        (pass_src, pass_dst), (base_src, base_dst) = pass_ids, bases
        ns, nt, R = bsrc.n, bstr.n, cfg.roi_batch
        n = ns + nt
        T = self._teacher_out
        pseudo, t_feats, t_lv, t_rpn_out = T["pseudo"], T["feats"], T["lv"], T["rpn_out"]
        images = self._mb_cache[("canvas", ns, nt, bsrc.hp, bsrc.wp)]
        sizes = torch.empty(n, 2, dtype=torch.int32, device=self.device)
        sizes[:ns].copy_(bsrc.sizes)
        sizes[ns:].copy_(bstr.sizes)
        feats, saved = det.backbone(W, images, sizes, save=True)
        lv = det.levels(feats)
        rpn_out, rpn_ts = det.rpn_head(W, feats, lv, save=True)
        props = det.proposals(rpn_out, lv, sizes, cfg.rpn_pre_topk[0], cfg.rpn_post_topk[0], cfg.rpn_nms_thresh, self.err_flag)
        props = self._override_proposals(props, [(pass_src, ns), (pass_dst, nt)])
        m = n * R
        rois = torch.empty(m, 4, device=self.device)
        roi_gt = torch.empty(m, 4, device=self.device)
        roi_batch = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_class = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_src = torch.empty(m, dtype=torch.int32, device=self.device)
        roi_count = torch.zeros(n, dtype=torch.int32, device=self.device)
        roi_stats = torch.zeros(n, 2, dtype=torch.int32, device=self.device)
        pstride = props["boxes"].shape[1]
        halves = ((bsrc, bsrc.gt, 0, ns), (bstr, pseudo, ns, nt))
        for b, gt, i0, k in halves:
            r0, r1 = i0 * R, (i0 + k) * R
            ops.call("aldi_roi_label_sample", props["boxes"][i0:i0 + k], props["count"][i0:i0 + k], pstride, k, gt.boxes,
                     gt.classes, gt.counts, gt.gmax, cfg.roi_iou, cfg.num_classes, R, cfg.roi_pos_fraction, b.seed,
                     b.salts[sampling.SITE_ROI], 1, rois[r0:r1], roi_batch[r0:r1], roi_class[r0:r1], roi_gt[r0:r1],
                     roi_src[r0:r1], roi_count[i0:i0 + k], roi_stats[i0:i0 + k])
        groups = [(0, ns * R, 0, ns), (ns * R, nt * R, ns, nt)]
        pred, head_saved = det.box_head(W, feats, rois, roi_batch, save=True, groups=groups)
        # teacher RoI head on the student's sampled TARGET proposals (ReplaceProposalsOnce + shared seed)
        t_pred, _ = det.box_head(self.teacher, t_feats, rois[ns * R:], roi_batch[ns * R:], save=False)
        d_rpn = torch.zeros(n, lv.total_locs, 64, device=self.device, dtype=self.dtype)
        dpred = torch.zeros(m, 64, device=self.device, dtype=self.dtype)
        w10 = ops.host_floats((10.0, 10.0, 5.0, 5.0))
        # ---- source half: hard losses (aldi/trainer.py:86-90)
        labels, matched, _ = self._label_anchors(lv, bsrc, bsrc.gt, sampling.SITE_RPN)
        ops.call("aldi_rpn_loss", rpn_out[:ns], _l.ctypes.byref(lv), ns, labels, matched, bsrc.gt.boxes, bsrc.gt.counts,
                 bsrc.gt.gmax, cfg.rpn_batch, 1.0, 1.0, gscale, d_rpn[:ns], self.dtc, 64, 0,
                 self.loss_acc[base_src + 2:base_src + 4])
        ops.call("aldi_roi_loss", pred[:ns * R], det.PRED_CH, ns * R, cfg.num_classes, roi_class[:ns * R], rois[:ns * R],
                 roi_gt[:ns * R], roi_count[:ns], ns, w10, 1.0, 1.0, gscale, dpred[:ns * R], self.dtc, 64,
                 self.loss_acc[base_src:base_src + 2])
        # ---- target half: distillation losses (aldi/distill.py:170-278)
        hard_rpn = cfg.do_hard_obj or cfg.do_hard_rpn_reg
        acc = 0
        if hard_rpn:
            hl, hm, _ = self._label_anchors(lv, bstr, pseudo, sampling.SITE_RPN)
            ops.call("aldi_rpn_loss", rpn_out[ns:], _l.ctypes.byref(lv), nt, hl, hm, pseudo.boxes, pseudo.counts, pseudo.gmax,
                     cfg.rpn_batch, 1.0 if cfg.do_hard_obj else 0.0, 1.0 if cfg.do_hard_rpn_reg else 0.0, gscale,
                     d_rpn[ns:], self.dtc, 64, 0, self.loss_acc[base_dst + 2:base_dst + 4])
            acc = 1
        dlabels, _, dstats = self._label_anchors(t_lv, bstr, pseudo, sampling.SITE_RPN_DISTILL)   # T2
        ops.call("aldi_distill_rpn_loss", rpn_out[ns:], t_rpn_out, _l.ctypes.byref(lv), nt, dlabels, dstats,
                 cfg.obj_temperature, 1.0 if cfg.do_obj_dst else 0.0, 1.0 if cfg.do_rpn_reg_dst else 0.0, gscale, d_rpn[ns:],
                 self.dtc, 64, acc, self.loss_acc[base_dst + 4:base_dst + 6])
        acc = 0
        if cfg.do_hard_cls or cfg.do_hard_roi_reg:
            ops.call("aldi_roi_loss", pred[ns * R:], det.PRED_CH, nt * R, cfg.num_classes, roi_class[ns * R:], rois[ns * R:],
                     roi_gt[ns * R:], roi_count[ns:], nt, w10, 1.0 if cfg.do_hard_cls else 0.0,
                     1.0 if cfg.do_hard_roi_reg else 0.0, gscale, dpred[ns * R:], self.dtc, 64,
                     self.loss_acc[base_dst:base_dst + 2])
            acc = 1
        ops.call("aldi_distill_roi_loss", pred[ns * R:], t_pred, det.PRED_CH, nt * R, cfg.num_classes, roi_class[ns * R:],
                 roi_count[ns:], nt, cfg.cls_temperature, 1 if cfg.cls_loss_type == "KL" else 0,
                 1.0 if cfg.do_cls_dst else 0.0, 1.0 if cfg.do_roih_reg_dst else 0.0, gscale, dpred[ns * R:], self.dtc, 64, acc,
                 self.loss_acc[base_dst + 6:base_dst + 8])
        if self.debug is not None:
            fw_t = {"lv": lv, "rpn_out": rpn_out[ns:], "pred": pred[ns * R:], "roi_count": roi_count[ns:],
                    "rois": rois[ns * R:], "roi_class": roi_class[ns * R:], "roi_batch": roi_batch[ns * R:]}
            self.debug = {"fw": fw_t, "t_pred": t_pred, "t_rpn_out": t_rpn_out, "labels": dlabels, "stats": dstats,
                          "pseudo": pseudo}
        det.backward(W, self.grad, feats, saved, rpn_ts, d_rpn, lv, head_saved, dpred, rois, roi_batch,
                     on_ready=self._bucket_ready(), **self._bwd_kw(), groups=groups)

    # ---- one distillation micro-batch (aldi/distill.py:144-278) ------------------------------------------
    def _distill_body(self, bw, bs, gscale, pass_id, base=SLOT_BASE["distill"]):
        cfg = self.cfg
        ctx = self._distill_forward(bw, bs, pass_id)
        w = {"loss_cls": gscale if cfg.do_hard_cls else 0.0, "loss_box_reg": gscale if cfg.do_hard_roi_reg else 0.0,
             "loss_rpn_cls": gscale if cfg.do_hard_obj else 0.0, "loss_rpn_loc": gscale if cfg.do_hard_rpn_reg else 0.0,
             "loss_obj_bce": gscale if cfg.do_obj_dst else 0.0, "loss_rpn_l1": gscale if cfg.do_rpn_reg_dst else 0.0,
             "loss_cls_ce": gscale if cfg.do_cls_dst else 0.0, "loss_roih_l1": gscale if cfg.do_roih_reg_dst else 0.0}
        self._distill_finish(ctx, w, self.loss_acc, base)

    def _distill_forward(self, bw, bs, pass_id):
        """aldi/distill.py:144-168: pseudo-label with the teacher, student forward on the strong views, teacher box
        head on the student's sampled proposals, fresh teacher anchor sampling on the pseudo labels (T2)."""
        cfg, det = self.cfg, self.det
        t_feats, t_lv, t_rpn_out = self.teacher_forward(bw)
        pseudo, _ = self.pseudo_label(bw, t_feats, t_lv, t_rpn_out)
        self.last_pseudo = pseudo
        if self.debug is not None:
            self.pseudo_log.append(pseudo)
        if self.pseudo_override is not None:
            pseudo = self.pseudo_override[len(self.pseudo_log) - 1]
        hard_rpn = cfg.do_hard_obj or cfg.do_hard_rpn_reg
        fw = self._student_forward(bs, pseudo, pass_id, want_rpn_labels=hard_rpn)
        if self.teacher.bottom_up is not None and any(r > 0 for r in self.teacher.bottom_up.drop_rates):
            # stochastic bottom-up: the soft targets come from a separate TRAINING-mode teacher forward
            t_feats, t_lv, t_rpn_out = self.teacher_forward(bw, train_mode=True)
        # teacher RoI head on the student's sampled proposals (ReplaceProposalsOnce + shared seed)
        t_pred, _ = det.box_head(self.teacher, t_feats, fw["rois"], fw["roi_batch"], save=False)
        # fresh anchor sampling by the TEACHER's RPN on the pseudo labels (T2)
        labels, _, stats = self._label_anchors(t_lv, bs, pseudo, sampling.SITE_RPN_DISTILL)
        self.debug = {"fw": fw, "t_pred": t_pred, "t_rpn_out": t_rpn_out, "labels": labels, "stats": stats,
                      "pseudo": pseudo} if self.debug is not None else None
        return {"fw": fw, "t_pred": t_pred, "t_rpn_out": t_rpn_out, "labels": labels, "stats": stats, "pseudo": pseudo,
                "n": bw.n}

    def _distill_finish(self, ctx, w, loss_vec, base):
        """Losses of aldi/distill.py:170-278 with one weight per key (0 = the reference's `* 0.0`, T5) and the backward."""
        fw = ctx["fw"]
        d_rpn, dpred = self._distill_losses(ctx, w, loss_vec, base)
        self.det.backward(self.student, self.grad, fw["feats"], fw["saved"], fw["rpn_ts"], d_rpn, fw["lv"], fw["head_saved"],
                          dpred, fw["rois"], fw["roi_batch"], on_ready=self._bucket_ready(), **self._bwd_kw())

    def _distill_losses(self, ctx, w, loss_vec, base):
        cfg, det = self.cfg, self.det
        fw, pseudo, n = ctx["fw"], ctx["pseudo"], ctx["n"]
        lv = fw["lv"]
        d_rpn = torch.zeros(n, lv.total_locs, 64, device=self.device, dtype=self.dtype)
        acc = 0
        if (w["loss_rpn_cls"] or w["loss_rpn_loc"]) and "labels" in fw:
            ops.call("aldi_rpn_loss", fw["rpn_out"], _l.ctypes.byref(lv), n, fw["labels"], fw["matched"], pseudo.boxes,
                     pseudo.counts, pseudo.gmax, cfg.rpn_batch, w["loss_rpn_cls"], w["loss_rpn_loc"], 1.0, d_rpn, self.dtc,
                     64, 0, loss_vec[base + 2:base + 4])
            acc = 1
        ops.call("aldi_distill_rpn_loss", fw["rpn_out"], ctx["t_rpn_out"], _l.ctypes.byref(lv), n, ctx["labels"], ctx["stats"],
                 cfg.obj_temperature, w["loss_obj_bce"], w["loss_rpn_l1"], 1.0, d_rpn, self.dtc, 64, acc,
                 loss_vec[base + 4:base + 6])
        m = n * cfg.roi_batch
        dpred = torch.zeros(m, 64, device=self.device, dtype=self.dtype)
        acc = 0
        if w["loss_cls"] or w["loss_box_reg"]:
            ops.call("aldi_roi_loss", fw["pred"], det.PRED_CH, m, cfg.num_classes, fw["roi_class"], fw["rois"],
                     fw["roi_gt"], fw["roi_count"], n, ops.host_floats((10.0, 10.0, 5.0, 5.0)), w["loss_cls"],
                     w["loss_box_reg"], 1.0, dpred, self.dtc, 64, loss_vec[base:base + 2])
            acc = 1
        ops.call("aldi_distill_roi_loss", fw["pred"], ctx["t_pred"], det.PRED_CH, m, cfg.num_classes, fw["roi_class"],
                 fw["roi_count"], n, cfg.cls_temperature, 1 if cfg.cls_loss_type == "KL" else 0, w["loss_cls_ce"],
                 w["loss_roih_l1"], 1.0, dpred, self.dtc, 64, acc, loss_vec[base + 6:base + 8])
        return d_rpn, dpred

    # ---- module-facade seams (aldi_b200/model.py): ONE micro-batch per call, losses first, backward when autograd asks ----
    def _facade_begin(self):
        if not getattr(self, "_facade_open", False):
            self._facade_open, self._facade_pass = True, 0
            self.seed = random.randint(0, 2 ** 32 - 1)      # aldi/helpers.py:19-23, as run_model draws it
            self.seed_log = {}
            self.loss_acc.zero_()
            self._last_backward = False

    def facade_forward(self, data, labeled=True, do_align=False):
        """`model(batched_inputs, labeled=, do_align=)` of the reference (aldi/model.py:27-29) for one micro-batch:
        student forward + loss values now, `finish(weights)` runs the loss gradients and the explicit backward later.
        -> (keys, values: fp32 device vector of the UNWEIGHTED losses, finish)"""
        self._facade_begin()
        pass_id = self._facade_pass
        self._facade_pass += 1
        self.seed_log[pass_id] = self.seed
        kind = "source" if labeled else "target"
        item = {"keys": [(kind,) + MicroBatch.shape_key(data, labeled)], "parts": [(kind, data, labeled)],
                "seed": self.seed, "pass_id": pass_id}
        b, = self._stage(item, None)
        fw = self._student_forward(b, b.gt, pass_id, want_rpn_labels=labeled)
        if getattr(self, "_scratch_grad", None) is None and do_align:
            self._scratch_grad = torch.zeros_like(self.grad)
        vals = torch.zeros(8, device=self.device)
        ones = {k: 1.0 for k in SRC_KEYS + ("da",)}
        # unlabeled (target_weak) passes carry no detection losses: their slots stay 0, the alignment pair lands at 4..5
        self._source_losses(fw, b, ones, vals, 0 if labeled else 4, labeled, do_align, getattr(self, "_scratch_grad", None))
        keys = list(SRC_KEYS) + ([k for k, on in zip(DA_KEYS, (self.cfg.img_da_enabled, self.cfg.ins_da_enabled)) if on]
                                 if do_align else [])
        idx = list(range(4)) + ([4 + j for j, on in enumerate((self.cfg.img_da_enabled, self.cfg.ins_da_enabled)) if on]
                                if do_align else [])

        def finish(weights):
            w = {k: float(weights.get(k, 0.0)) for k in SRC_KEYS}
            da = [float(weights[k]) for k in DA_KEYS if k in weights]
            assert len(set(da)) <= 1, "the alignment losses share one weight (aldi/trainer.py:75-79)"
            w["da"] = da[0] if da else 0.0
            self._source_finish(fw, b, w, torch.zeros(8, device=self.device), 0, labeled=labeled, do_align=do_align)
            b.free_event = torch.cuda.Event()
            b.free_event.record()

        return keys, vals[idx], finish

    def facade_distill(self, teacher_data, student_data):
        """`distiller(teacher_batched_inputs, student_batched_inputs)` (aldi/distill.py:170-191) for one micro-batch.
        -> (keys: the 4 hard + the enabled soft loss names, values, finish)"""
        self._facade_begin()
        cfg = self.cfg
        pass_id = 100 + self._facade_pass
        self._facade_pass += 1
        seed = random.randint(0, 2 ** 32 - 1)               # seeder.reset_seed(), aldi/distill.py:150
        self.seed_log[pass_id] = seed
        self.seed = seed
        item = {"keys": [("weak",) + MicroBatch.shape_key(teacher_data, False), ("strong",) + MicroBatch.shape_key(student_data, False)],
                "parts": [("weak", teacher_data, False), ("strong", student_data, False)], "seed": seed, "pass_id": pass_id}
        bw, bs = self._stage(item, None)
        ctx = self._distill_forward(bw, bs, pass_id)
        on = {"loss_cls": cfg.do_hard_cls, "loss_box_reg": cfg.do_hard_roi_reg, "loss_rpn_cls": cfg.do_hard_obj,
              "loss_rpn_loc": cfg.do_hard_rpn_reg, "loss_obj_bce": cfg.do_obj_dst, "loss_rpn_l1": cfg.do_rpn_reg_dst,
              "loss_cls_ce": cfg.do_cls_dst, "loss_roih_l1": cfg.do_roih_reg_dst}
        vals = torch.zeros(8, device=self.device)
        self._distill_losses(ctx, {k: 1.0 if v else 0.0 for k, v in on.items()}, vals, 0)
        keys = list(SRC_KEYS) + [k for k in SOFT_KEYS if on[k]]
        idx = list(range(4)) + [4 + j for j, k in enumerate(SOFT_KEYS) if on[k]]

        def finish(weights):
            self._distill_finish(ctx, {k: float(weights.get(k, 0.0)) if on[k] else 0.0 for k in on},
                                 torch.zeros(8, device=self.device), 0)
            for m in (bw, bs):
                m.free_event = torch.cuda.Event()
                m.free_event.record()

        return keys, vals[idx], finish

    # ---- aldi/dropin.py:121 optimizer.step() (torch.optim.SGD via D2 build_optimizer) ---------------------
    def lr_at(self, it, warmup_iters=100, warmup_factor=0.01, steps=(), gamma=0.1):
        """detectron2 WarmupMultiStepLR (linear warm-up), configs/Base-RCNN-FPN.yaml:12-21."""
        lr = self.cfg.base_lr * gamma ** sum(1 for s in steps if it >= s)
        if it < warmup_iters:
            a = it / warmup_iters
            lr *= warmup_factor * (1 - a) + a
        return lr

    def _bwd_kw(self):
        # sparse RPN backward: a strict bound on the locations with a non-zero head gradient per image -- 256 sampled
        # anchors of the hard loss + 256 valid entries and 128 x 4 flat delta entries of the distillation loss (T1)
        return {"rpn_rows_per_image": 4 * self.cfg.rpn_batch, "err_flag": self.err_flag}

    def _bucket_ready(self):
        """Gradient buckets become final only in the step's LAST backward; earlier micro-batches just accumulate."""
        if self._capture_hook is not None:
            return self._capture_hook
        if not (self._last_backward and self.reducer.active):
            return None
        return self.reducer.ready

    def allreduce_grads(self):
        """One sum-all-reduce of the flat gradient buffer per step, started bucket by bucket during the last
        backward (data_parallel.GradReducer); DDP in the reference averages inside every micro-batch backward."""
        return self.reducer.finish()

    def optimizer_step(self, lr=None):
        lr = self.lr_at(self.iter) if lr is None else lr
        gs = self.allreduce_grads()
        p = self.student.flat[:self.nt]
        bu = self.student.bottom_up
        if bu is not None and self.reducer.active:
            dist.all_reduce(bu.grad, op=dist.ReduceOp.SUM, group=self.pg)
        if self.cfg.optimizer.upper() == "ADAMW":
            b1, b2 = self.cfg.adamw_betas
            ops.call("aldi_adamw_step", p, self.momentum_buf, self.exp_avg_sq, self.grad, self.nt, lr, b1, b2, self.cfg.adamw_eps,
                     self.cfg.weight_decay, self.iter + 1, gs)
            if bu is not None and hasattr(bu, "opt_segments"):
                # ViTDet (aldi/backbone.py:66-84): layer-wise lr decay, no decay on torch.nn.LayerNorm weights and pos_embed
                for off, n, factor, no_wd in bu.opt_segments():
                    ops.call("aldi_adamw_step", bu.flat[off:off + n], self.bu_exp_avg[off:off + n], self.bu_exp_avg_sq[off:off + n],
                             bu.grad[off:off + n], n, lr * factor, b1, b2, self.cfg.adamw_eps,
                             0.0 if no_wd else self.cfg.weight_decay, self.iter + 1, gs)
            elif bu is not None:
                ops.call("aldi_adamw_step", bu.flat, self.bu_exp_avg, self.bu_exp_avg_sq, bu.grad, bu.flat.numel(), lr, b1, b2,
                         self.cfg.adamw_eps, self.cfg.weight_decay, self.iter + 1, gs)
        else:
            ops.sgd_momentum_step(p, self.momentum_buf, self.grad, lr, self.cfg.weight_decay, self.cfg.momentum, gs)
            if bu is not None:
                ops.sgd_momentum_step(bu.flat, self.bu_momentum, bu.grad, lr, self.cfg.weight_decay, self.cfg.momentum, gs)
        self._facade_open = False
        self.grad.zero_()
        self.student.refresh(trainable_only=True)
        if bu is not None:
            bu.grad.zero_()
            bu.refresh()
        self.iter += 1

    # ---- a whole iteration: before_step + run_step ----------------------------------------------------------
    def step(self, data, lr=None):
        self.ema_update(self.iter)
        losses = self.run_model(data)
        self.optimizer_step(lr)
        return losses

    def state_dict(self, which="student"):
        w = self.student if which == "student" else self.teacher
        out = self.layout.unpack_state_dict(w.flat)
        if w.bottom_up is not None:
            pre = getattr(w.bottom_up, "prefix", "backbone.bottom_up.")
            out.update({pre + k: v for k, v in w.bottom_up.state_dict().items()})
        return out

    def load_state_dict(self, sd, which="student", strict=True):
        """Write a Detectron2-keyed state_dict into the flat buffer of the student or the teacher and re-derive the
        GEMM operands.  Returns (missing_keys, unexpected_keys) like nn.Module.load_state_dict; with strict=False keys
        that are absent keep their current values and tensors of the wrong shape are reported as missing."""
        w = self.student if which == "student" else self.teacher
        known = {key: (layer, field, off, n, shape) for (layer, field), (off, n, key, shape) in self.layout.entries.items()}
        bu_missing = []
        if w.bottom_up is not None:
            # the ConvNeXt bottom-up keeps its own flat buffer under the "backbone.bottom_up." prefix
            pre, lay = getattr(w.bottom_up, "prefix", "backbone.bottom_up."), w.bottom_up.layout
            keyed = getattr(w.bottom_up, "keyed_layout", False)
            host = w.bottom_up.flat.detach().cpu()
            for k, (off, n, shape) in lay.entries.items():
                t = sd.get(pre + k)
                if t is None or tuple(t.shape) != tuple(shape):
                    bu_missing.append(pre + k)
                else:
                    t = torch.as_tensor(t).detach().to("cpu", torch.float32)
                    host[off:off + n] = lay.to_internal(k, t) if keyed else lay.to_internal(t)
            w.bottom_up.flat.copy_(host)
            w.bottom_up.refresh()
            sd = {k: v for k, v in sd.items() if not (k.startswith(pre) and k[len(pre):] in lay.entries)}
        missing = [k for k in known if k not in sd] + bu_missing
        unexpected = [k for k in sd if k not in known]
        bad_shape = [k for k in known if k in sd and tuple(sd[k].shape) != tuple(known[k][4])]
        if strict and (missing or unexpected or bad_shape):
            raise KeyError("load_state_dict(strict): missing %s unexpected %s wrong shape %s"
                           % (missing[:5], unexpected[:5], bad_shape[:5]))
        host = w.flat.detach().cpu()
        for k, (layer, field, off, n, shape) in known.items():
            if k in sd and k not in bad_shape:
                host[off:off + n] = self.layout.to_internal(layer, field, torch.as_tensor(sd[k]).detach().to("cpu", torch.float32))
        w.flat.copy_(host)
        w.refresh()
        return missing + bad_shape, unexpected

    _OPT_BUFFERS = ("momentum_buf", "exp_avg_sq", "bu_momentum", "bu_exp_avg", "bu_exp_avg_sq")

    def optimizer_state(self):
        """Everything the optimizer needs to resume (the `optimizer` checkpointable of the reference trainer): the
        iteration and EVERY flat state buffer -- SGD momentum / AdamW first moment (`momentum_buf`), AdamW second moment,
        and the ConvNeXt bottom-up's own buffers.  `state` repeats the detector's buffers per parameter under the
        Detectron2 key names (torch.optim naming: momentum_buffer | exp_avg, exp_avg_sq) for tools that read those."""
        adamw = self.cfg.optimizer.upper() == "ADAMW"
        out = {"format": "aldi_b200.flat.v2", "optimizer": self.cfg.optimizer.upper(), "iteration": self.iter, "buffers": {}}
        for name in self._OPT_BUFFERS:
            t = getattr(self, name, None)
            if t is not None:
                out["buffers"][name] = t.detach().cpu().clone()
        names = {"momentum_buf": "exp_avg" if adamw else "momentum_buffer", "exp_avg_sq": "exp_avg_sq"}
        state = {}
        for buf, tname in names.items():
            t = out["buffers"].get(buf)
            if t is None:
                continue
            for (layer, field), (off, n, key, shape) in self.layout.entries.items():
                if off < self.nt and field in ("weight", "bias"):
                    state.setdefault(key, {})[tname] = self.layout.from_internal(layer, field, t[off:off + n], shape)
        out["state"] = state
        out["momentum_buffer"] = out["buffers"]["momentum_buf"]        # v1 readers
        return out

    def load_optimizer_state(self, state):
        """Accepts what `optimizer_state` writes (v2; v1 files with only `momentum_buffer` + `iteration` load what they
        hold).  A torch.optim state_dict (index-keyed `state` + `param_groups`, what the reference writes) cannot be
        mapped onto the flat buffers without the reference's parameter order: it is skipped with a warning and the
        optimizer restarts from zero moments at the stored iteration."""
        import logging
        log = logging.getLogger("aldi_b200")
        if not isinstance(state, dict) or ("param_groups" in state and "buffers" not in state):
            log.warning("optimizer state in torch.optim format: not loadable into the flat buffers, optimizer state reset")
            return
        bufs = dict(state.get("buffers") or {})
        if not bufs and "momentum_buffer" in state:
            bufs["momentum_buf"] = state["momentum_buffer"]
        if state.get("optimizer") and state["optimizer"] != self.cfg.optimizer.upper():
            log.warning("optimizer state was written by %s, this step runs %s: state reset", state["optimizer"],
                        self.cfg.optimizer.upper())
            bufs = {}
        for name in self._OPT_BUFFERS:
            dst = getattr(self, name, None)
            if dst is None:
                continue
            src = bufs.get(name)
            if src is None or src.numel() != dst.numel():
                if bufs:
                    log.warning("optimizer state: buffer %s missing or of another size in the checkpoint; zeroed", name)
                dst.zero_()
                continue
            dst.copy_(src.to(dst.device).reshape(dst.shape))
        self.iter = int(state.get("iteration", self.iter))


def detector_postprocess(results, output_height, output_width):
    """detectron2.modeling.postprocessing.detector_postprocess for box-only Instances: scale the boxes from the network
    input frame to the requested output resolution, clip to it, drop boxes that became empty."""
    from .structures import Boxes, Instances
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    b = results.pred_boxes.tensor.clone()
    b[:, 0::2] *= scale_x
    b[:, 1::2] *= scale_y
    b[:, 0].clamp_(min=0, max=output_width)
    b[:, 1].clamp_(min=0, max=output_height)
    b[:, 2].clamp_(min=0, max=output_width)
    b[:, 3].clamp_(min=0, max=output_height)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return Instances((output_height, output_width), pred_boxes=Boxes(b[keep]), scores=results.scores[keep],
                     pred_classes=results.pred_classes[keep])


class LossDict(dict):
    """dict[str -> float] materialised lazily with ONE device->host read (keys known without a sync)."""

    def __init__(self, step, keys):
        super().__init__()
        self._step, self._keys, self._done = step, keys, False
        for k, _ in keys:
            dict.__setitem__(self, k, None)

    def _materialise(self):
        if not self._done:
            for k, v in self._step.losses_to_host(self._keys).items():
                dict.__setitem__(self, k, v)
            self._done = True

    def __getitem__(self, k):
        self._materialise()
        return dict.__getitem__(self, k)

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)
