"""First measurement of BASELINE configs[4]: ALDI++ train step of Faster R-CNN on ConvNeXt-FPN at the CFC frame shape
(1920x1080 resized to 1024x1824 -> canvas 1024x1824 is already a multiple of 32), synthetic data, AdamW.

    python tools/bench_convnext.py [--size L|T] [--ims 2] [--steps 3]

Not the headline bench (that is bench.py on configs[1]); the ConvNeXt kernels are correctness-first.
Prints one JSON line: images/s over `steps` timed steps (CUDA events), peak memory, per-step ms.
"""
import argparse
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from aldi_b200 import arch, synth_data  # noqa: E402
from aldi_b200.convnext import synthetic_state_dict as convnext_init  # noqa: E402
from aldi_b200.train_step import B200TrainStep, StepConfig  # noqa: E402

SIZES = {"L": ((3, 3, 27, 3), (192, 384, 768, 1536)), "T": ((3, 3, 9, 3), (96, 192, 384, 768))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="L", choices=sorted(SIZES))
    ap.add_argument("--ims", type=int, default=2, help="source images = target images = SOLVER.IMS_PER_GPU")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--hw", default="1024x1824")
    ap.add_argument("--profile", action="store_true", help="also print device ms per entry point for one step (CUDA events)")
    args = ap.parse_args()
    h, w = (int(v) for v in args.hw.split("x"))
    depths, dims = SIZES[args.size]
    sd = arch.synthetic_state_dict(0, bottom_up_channels=dims)
    sd.update({"backbone.bottom_up." + k: v for k, v in convnext_init(depths, dims, 0, 1e-6).items()})
    cfg = StepConfig(dtype="bf16", ims_per_gpu=args.ims, backbone="convnext", convnext_depths=depths, convnext_dims=dims,
                     convnext_drop_path=0.2, optimizer="ADAMW", base_lr=1e-5, weight_decay=0.05,
                     anchor_sizes=((64,), (128,), (256,), (512,), (1024,)), pixel_std=(57.375, 57.12, 58.395))
    step = B200TrainStep(cfg, sd)
    step.debug = None
    ls, uw, us = synth_data.synthetic_batch(1234, args.ims, args.ims, h, w, num_boxes=12)
    dev = [[{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for d in b] for b in (ls, uw, us)]
    random.seed(0)
    for _ in range(2):
        step.step((None, dev[0], dev[1], dev[2]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = step.step((None, dev[0], dev[1], dev[2]))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    vals = dict(losses.items())
    if args.profile:
        from aldi_b200 import ops
        prof = ops.KernelProfiler()
        ops.set_profiler(prof)
        step.step((None, dev[0], dev[1], dev[2]))
        ops.set_profiler(None)
        for name, d in sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"])[:14]:
            print("%-28s %4d launches %8.2f ms" % (name, d["launches"], d["ms"]), file=sys.stderr)
    print(json.dumps({"workload": "ALDI++ Faster R-CNN ConvNeXt-%s FPN, %dx%d synthetic, %d source + %d target images, AdamW, "
                      "DropPath 0.2, eager (no CUDA graph)" % (args.size, h, w, args.ims, args.ims),
                      "images_per_s": 2 * args.ims / (ms / 1e3), "ms_per_step": ms, "steps": args.steps, "dtype": "bf16",
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                      "finite": all(v == v for v in vals.values()), "n_gpus": 1}))


if __name__ == "__main__":
    main()
