"""Host-side mirror of the reference's trainer for the path this repository covers.

`ALDITrainer` keeps the seams of `aldi.trainer.ALDITrainer` (aldi/trainer.py:137-246) — `build_train_loader`,
`before_step` (EMA update), `run_step` (`run_model` + metrics + optimizer step), `train(start_iter, max_iter)` —
but every seam lands in `B200TrainStep`.  It exists so that a cfg built from the reference's own YAML keys drives
the B200 step end to end without Detectron2 (which cannot be installed offline); with Detectron2 present the
reference's trainer is subclassed instead.  Control-plane features of the reference trainer
(hooks, periodic evaluation, checkpoint rotation, TensorBoard writers) are out of scope.
"""
import logging
import time

import torch
import torch.distributed as dist

from . import arch, synth_data
from .config import step_config_from_cfg
from .data_parallel import reduce_loss_vector
from .model import EMA, build_aldi, build_distiller

logger = logging.getLogger("aldi_b200")


def unpack_data_weak_strong(labeled, unlabeled, batch_contents):
    """-> (labeled_weak, labeled_strong, unlabeled_weak, unlabeled_strong) following aldi/dataloader.py:57-80:
    a weak view is returned when asked for, and the unlabeled weak view ALSO whenever any unlabeled data is
    requested (pseudo-labelling needs it).  Items carry the weak image under "img_weak" when the strong view
    replaced "image" (the reference's WEAK_IMG_KEY)."""

    def weak(batch):
        out = []
        for d in batch:
            e = dict(d)
            e.pop("aug_params", None)          # the weak view is what the strong one is derived FROM
            if "img_weak" in e:
                e["image"] = e["img_weak"]
            out.append(e)
        return out

    labeled_weak = weak(labeled) if "labeled_weak" in batch_contents and labeled is not None else None
    labeled_strong = labeled if "labeled_strong" in batch_contents else None
    unlabeled_weak = None
    if ("unlabeled_weak" in batch_contents or "unlabeled_strong" in batch_contents) and unlabeled is not None:
        unlabeled_weak = weak(unlabeled)
    unlabeled_strong = unlabeled if "unlabeled_strong" in batch_contents else None
    return labeled_weak, labeled_strong, unlabeled_weak, unlabeled_strong


class SyntheticWeakStrongLoader:
    """Infinite iterator of 4-tuples in the format of aldi.dataloader.WeakStrongDataloader over synthetic COCO-style
    images (there is no dataset on the box): per-rank batch sizes follow ALDITrainer.build_train_loader
    (aldi/trainer.py:211-240): IMS_PER_BATCH split by BATCH_RATIOS, divided by the world size."""

    def __init__(self, cfg, height, width, rank=0, world=1, seed=1234, pin=False, gpu_aug=False):
        contents, ratios = cfg.DATASETS.BATCH_CONTENTS, cfg.DATASETS.BATCH_RATIOS
        total = cfg.SOLVER.IMS_PER_BATCH
        sizes = [int(total * r / sum(ratios)) for r in ratios]
        assert len(contents) == len(sizes), "len(cfg.DATASETS.BATCH_CONTENTS) must equal len(cfg.DATASETS.BATCH_RATIOS)."
        assert sum(sizes) == total, "sum(batch_sizes)=%d must equal total_batch_size=%d" % (sum(sizes), total)
        lab = [s for s, c in zip(sizes, contents) if c.startswith("labeled")]
        unl = [s for s, c in zip(sizes, contents) if c.startswith("unlabeled")]
        self.labeled_bs = (max(lab) if lab else 0) // world
        self.unlabeled_bs = (max(unl) if unl else 0) // world
        self.contents, self.h, self.w, self.pin = contents, height, width, pin
        self.seed, self.rank, self.it = seed, rank, 0
        self.num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        # gpu_aug: ship only the weak view; strong items carry the drawn augmentation parameters and the step derives
        # the strong view in HBM (aldi_b200/augment.py) — what get_augs(cfg, labeled, ...) would do on the workers
        self.augmenters = None
        if gpu_aug:
            from .augment import StrongAugmenter
            self.augmenters = {"labeled": StrongAugmenter.from_config(cfg, True),
                               "unlabeled": StrongAugmenter.from_config(cfg, False)}

    def __iter__(self):
        return self

    def __next__(self):
        ls, uw, us = synth_data.synthetic_batch(self.seed + 7919 * self.it + self.rank, self.labeled_bs, self.unlabeled_bs,
                                                self.h, self.w, num_classes=self.num_classes)
        self.it += 1
        labeled = ls if self.labeled_bs else None
        unlabeled = None
        if self.unlabeled_bs:
            unlabeled = [dict(s, img_weak=w["image"]) for w, s in zip(uw, us)]
        if self.augmenters is not None:
            # the synthetic generator's own strong views are dropped: labeled items keep the weak image + parameters,
            # unlabeled strong items are image-less (their weak twin is staged right before them)
            if labeled is not None:
                labeled = [dict(d, aug_params=self.augmenters["labeled"].draw(self.h, self.w)) for d in labeled]
            if unlabeled is not None:
                unlabeled = [{"image": None, "img_weak": d["img_weak"], "height": d["height"], "width": d["width"],
                              "aug_params": self.augmenters["unlabeled"].draw(self.h, self.w)} for d in unlabeled]
        if self.pin:
            for b in (labeled or []) + (unlabeled or []):
                for k in ("image", "img_weak"):
                    if k in b:
                        b[k] = b[k].pin_memory()
        return unpack_data_weak_strong(labeled, unlabeled, self.contents)


class ALDITrainer:
    def __init__(self, cfg, state_dict=None, data_loader=None, process_group=None, device=None, dtype=None,
                 image_size=(512, 512)):
        self.cfg = cfg
        self.pg = process_group
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        if cfg.MODEL.DEVICE != "cuda" and device is None:
            raise RuntimeError("aldi_b200 runs the train step on a CUDA device only (MODEL.DEVICE=%s): there is no CPU "
                               "fallback; the CPU arm of bench.py is the oracle" % cfg.MODEL.DEVICE)
        scfg = step_config_from_cfg(cfg, dtype=dtype)
        sd = state_dict
        if sd is None:
            from .train_step import synthetic_state_dict_for
            sd = synthetic_state_dict_for(scfg, 0, cfg.MODEL.CONVNEXT.LAYER_SCALE_INIT_VALUE)
        # aldi/trainer.py:139-147,156-160: the model, its EMA teacher and the distiller come from the registries named in cfg
        # (MODEL.META_ARCHITECTURE, DOMAIN_ADAPT.ALIGN.MIXIN_NAME, DOMAIN_ADAPT.DISTILL.MIXIN_NAME / DISTILLER_NAME); all
        # three are facades over ONE engine, which is what run_step drives
        self.model = build_aldi(cfg, state_dict=sd, device=device or "cuda:%d" % torch.cuda.current_device(),
                                process_group=process_group, dtype=dtype)
        self.step_impl = self.model.engine
        self.ema = EMA(self.model, cfg.EMA.ALPHA, cfg.EMA.START_ITER) if cfg.EMA.ENABLED else None
        if self.ema is None:
            # aldi/trainer.py:142: without EMA the "teacher" handed to the distiller is the student itself
            self.step_impl.teacher = self.step_impl.student
        self.distiller = build_distiller(cfg, teacher=self.ema.model if self.ema is not None else self.model,
                                         student=self.model)
        self.data_loader = data_loader if data_loader is not None else self.build_train_loader(cfg, image_size, self.rank,
                                                                                               self.world)
        if getattr(self.data_loader, "augmenters", None):
            self.step_impl.augmenters = self.data_loader.augmenters
        self._data_iter = iter(self.data_loader)
        self.iter = self.start_iter = 0
        self.max_iter = cfg.SOLVER.MAX_ITER
        self.history = []

    # ---- aldi/trainer.py:211-240 ---------------------------------------------------------------------------
    @classmethod
    def build_train_loader(cls, cfg, image_size=(512, 512), rank=0, world=1, gpu_aug=False):
        return SyntheticWeakStrongLoader(cfg, image_size[0], image_size[1], rank=rank, world=world, pin=True, gpu_aug=gpu_aug)

    # ---- D2 WarmupMultiStepLR via SOLVER.* --------------------------------------------------------------------
    def lr(self, it):
        s = self.cfg.SOLVER
        return self.step_impl.lr_at(it, warmup_iters=s.WARMUP_ITERS, warmup_factor=s.WARMUP_FACTOR, steps=tuple(s.STEPS),
                                    gamma=s.GAMMA)

    # ---- aldi/trainer.py:242-246 ----------------------------------------------------------------------------
    def before_step(self):
        if self.cfg.EMA.ENABLED:
            self.ema.update_weights(self.model, self.iter)

    # ---- aldi/dropin.py:94-121 (SimpleTrainer.run_step with the run_model / do_backward seams) -------------------
    def run_step(self):
        t0 = time.perf_counter()
        data = next(self._data_iter)
        data_time = time.perf_counter() - t0
        loss_dict = self.step_impl.run_model(data)          # forward + backward of every micro-batch
        self.step_impl.optimizer_step(self.lr(self.iter))   # all-reduce wait + fused SGD + operand refresh
        self._write_metrics(loss_dict, data_time)

    def _write_metrics(self, loss_dict, data_time):
        """One loss-vector read-back per iteration, averaged over ranks (aldi/dropin.py:100-118 gathers dicts)."""
        vals = dict(loss_dict.items())
        if self.world > 1:
            keys = sorted(vals)
            v = reduce_loss_vector(torch.tensor([vals[k] for k in keys], device=self.step_impl.device), self.pg)
            vals = dict(zip(keys, v.tolist()))
        total = sum(vals.values())
        if not all(x == x and abs(x) != float("inf") for x in vals.values()):
            raise FloatingPointError("Loss became infinite or NaN at iteration=%d!\nloss_dict = %s" % (self.iter, vals))
        self.history.append(dict(vals, total_loss=total, data_time=data_time, lr=self.lr(self.iter)))

    def train(self, start_iter=0, max_iter=None):
        self.iter = self.start_iter = start_iter
        self.max_iter = max_iter if max_iter is not None else self.max_iter
        for self.iter in range(start_iter, self.max_iter):
            self.before_step()
            self.run_step()
            if self.rank == 0 and (self.iter % 20 == 0 or self.iter == self.max_iter - 1):
                h = self.history[-1]
                logger.info("iter %d  total_loss %.4f  lr %.6f  %s", self.iter, h["total_loss"], h["lr"],
                            "  ".join("%s %.4f" % (k, v) for k, v in h.items() if k.startswith("loss_")))
        self.iter += 1
        return self.history

    def state_dict(self):
        """{"model", "ema"} with Detectron2 key names, the layout DetectionCheckpointerWithEMA reads
        (aldi/checkpoint.py:8-31)."""
        out = {"model": self.step_impl.state_dict("student"), "iteration": self.iter}
        if self.cfg.EMA.ENABLED:
            out["ema"] = self.step_impl.state_dict("teacher")
        return out
