"""Micro-benchmark of the ConvNeXt stream kernels (csrc/convnext.cu) at the ConvNeXt-L stage geometries of a 1024 x 1824
input (SURVEY §8 a18, configs[4]): CUDA-event medians with an L2 flush between iterations, algorithmic bytes (read each
input once, write each output once) against the measured HBM copy peak, and -- for the depthwise stencil -- the fp32
FMA issue floor (49 FMA per output element; 148 SMs x 128 lanes).

    python tools/bench_convnext_ops.py [--n 2] [--iters 10] [--only dwconv7,gelu]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from aldi_b200 import ops  # noqa: E402

STAGES = [(256, 456, 192), (128, 228, 384), (64, 114, 768), (32, 57, 1536)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--stages", default="0,1,2,3")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    peak = 6544.7
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sm_clock_ghz = 1.9

    def timed(fn):
        ts = []
        for _ in range(args.iters + 3):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[3:])
        return ts[len(ts) // 2]

    rows_out = []
    n = args.n
    g = torch.Generator(device="cuda").manual_seed(0)
    for h, w, c in [STAGES[int(i)] for i in args.stages.split(",")]:
        rows = n * h * w
        x = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
        dy = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
        y = torch.empty_like(x)
        wt = torch.randn(c, 49, device="cuda", generator=g) * 0.1
        b = torch.randn(c, device="cuda", generator=g)
        dw = torch.zeros(c, 49, device="cuda")
        gamma, beta = torch.rand(c, device="cuda", generator=g) + 0.5, torch.randn(c, device="cuda", generator=g)
        stats = torch.empty(rows, 2, device="cuda")
        dgam, dbet = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        hid = torch.randn(rows, 4 * c, device="cuda", generator=g).bfloat16()
        da = torch.randn(rows, 4 * c, device="cuda", generator=g).bfloat16()
        act = torch.empty_like(hid)
        t = x.numel() * 2 / 1e9                                   # GB of one activation tensor
        cases = {
            "dwconv7": (lambda: ops.call("aldi_dwconv7", x, wt, b, n, h, w, c, c, 1, 0, y, 0), 2 * t, rows * c * 49),
            "dwconv7_wgrad": (lambda: ops.call("aldi_dwconv7_wgrad", x, dy, n, h, w, c, c, 1, dw), 2 * t, rows * c * 49),
            "layernorm_forward": (lambda: ops.call("aldi_layernorm_forward", x, gamma, beta, 1e-6, rows, c, c, 1, y, stats), 2 * t, 0),
            "layernorm_backward": (lambda: ops.call("aldi_layernorm_backward", x, gamma, stats, dy, rows, c, c, 1, y, 0, dgam, dbet),
                                   3 * t, 0),
            "layerscale_forward": (lambda: ops.call("aldi_layerscale_forward", x, dy, gamma, None, rows, h * w, c, c, 1, y), 3 * t, 0),
            "gelu": (lambda: ops.call("aldi_gelu", hid, None, act, hid.numel(), 1), 8 * t, 0),
            "gelu_backward": (lambda: ops.call("aldi_gelu", hid, da, act, hid.numel(), 1), 12 * t, 0),
        }
        ops.call("aldi_layernorm_forward", x, gamma, beta, 1e-6, rows, c, c, 1, y, stats)
        for name, (fn, gb, fma) in cases.items():
            if only and name not in only:
                continue
            ms = timed(fn)
            rec = {"op": name, "n": n, "h": h, "w": w, "c": c, "ms": round(ms, 4), "algorithmic_gb": round(gb, 4),
                   "gbs": round(gb / ms * 1e3, 1), "hbm_frac": round(gb / ms * 1e3 / peak, 3)}
            if fma:
                floor_ms = fma / (148 * 128 * sm_clock_ghz * 1e9) * 1e3
                rec["fma_floor_ms"] = round(floor_ms, 4)
                rec["fma_frac"] = round(floor_ms / ms, 3)
            rows_out.append(rec)
            print(json.dumps(rec), flush=True)
        del x, dy, y, hid, da, act
    return rows_out


if __name__ == "__main__":
    main()
