#!/usr/bin/env python
"""Command line of the B200 train step with the reference's interface (tools/train_net.py of justinkay/aldi):

    python tools/train_net.py --config-file <any ALDI R-CNN YAML, e.g. the reference's configs/cityscapes/ALDI-Best-Cityscapes.yaml> \
        [--num-gpus N] [--image-size H W] [--iters K] [KEY VALUE ...]

The YAML chain (`_BASE_`) and the `KEY VALUE` overrides are merged into `get_cfg()` + `add_aldi_config()` exactly
as the reference does; datasets are synthetic (no dataset / checkpoint is reachable offline), weights are the
deterministic synthetic initialisation unless MODEL.WEIGHTS points at a torch-saved Detectron2 state dict.
Multi-GPU: launch under torchrun (one process per GPU); --num-gpus is checked against WORLD_SIZE.
"""
import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def setup(args):
    from aldi_b200.config import add_aldi_config, get_cfg
    cfg = get_cfg()
    add_aldi_config(cfg)
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    cfg.merge_from_list(args.opts)
    cfg.freeze()
    return cfg


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-file", default="", metavar="FILE")
    ap.add_argument("--num-gpus", type=int, default=1)
    ap.add_argument("--image-size", type=int, nargs=2, default=(512, 512), metavar=("H", "W"))
    ap.add_argument("--iters", type=int, default=None, help="override SOLVER.MAX_ITER")
    ap.add_argument("opts", nargs=argparse.REMAINDER, default=[])
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s %(name)s] %(message)s")
    import torch
    import torch.distributed as dist
    from aldi_b200.trainer import ALDITrainer
    cfg = setup(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.num_gpus, "--num-gpus %d but WORLD_SIZE=%d (launch with torchrun, one process per GPU)" % (
        args.num_gpus, world)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    trainer = ALDITrainer(cfg, process_group=pg, image_size=tuple(args.image_size),
                          dtype="bf16" if cfg.SOLVER.AMP.ENABLED else "fp32")
    if cfg.MODEL.WEIGHTS and os.path.isfile(cfg.MODEL.WEIGHTS):
        # DetectionCheckpointerWithEMA.resume_or_load(cfg.MODEL.WEIGHTS, resume=False): a .pth with an "ema" entry starts
        # the model from the EMA weights (aldi/checkpoint.py:19-31, tools/train_net.py:73-76 of the reference)
        from aldi_b200.checkpoint import DetectionCheckpointerWithEMA
        DetectionCheckpointerWithEMA(trainer.step_impl, cfg.OUTPUT_DIR).resume_or_load(cfg.MODEL.WEIGHTS, resume=False)
        trainer.step_impl.ema_update(-1)          # EMA(model) deep-copies the freshly loaded student (aldi/ema.py:12-13)
    elif cfg.MODEL.WEIGHTS:
        logging.getLogger("aldi_b200").warning("MODEL.WEIGHTS %s is not reachable here: synthetic initialisation", cfg.MODEL.WEIGHTS)
    hist = trainer.train(0, args.iters)
    if (not pg) or dist.get_rank() == 0:
        print({k: round(v, 5) for k, v in hist[-1].items()})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
