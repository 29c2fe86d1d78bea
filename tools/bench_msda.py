"""Micro-benchmark of the MSDeformAttn kernels at Deformable-DETR encoder geometry: CUDA-event timing,
algorithmic bytes (per head L*P*(4 taps*D + 3)*4 B read + D*4 B written) against the measured HBM peak.

    python tools/bench_msda.py [--n 2] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from aldi_b200.msda import MSDeformAttnFunction  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    shapes = [(100, 167), (50, 84), (25, 42), (13, 21)]   # 800 x 1333 input, strides 8..64
    starts, s = [], 0
    for h, w in shapes:
        starts.append(s)
        s += h * w
    n, m, d, p, l = args.n, 8, 32, 4, len(shapes)
    lq = s                                                   # encoder self-attention: every pixel is a query
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(n, s, m, d, device="cuda", generator=g, requires_grad=True)
    # encoder geometry: points within a few pixels of the query's own location
    ref = torch.rand(n, lq, 1, 1, 1, 2, device="cuda", generator=g)
    loc = (ref + 0.02 * torch.randn(n, lq, m, l, p, 2, device="cuda", generator=g)).requires_grad_(True)
    attn = torch.softmax(torch.randn(n, lq, m, l * p, device="cuda", generator=g), -1).view(n, lq, m, l, p).requires_grad_(True)
    sh, st = torch.as_tensor(shapes).cuda(), torch.as_tensor(starts).cuda()
    go = torch.randn(n, lq, m * d, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        ts = []
        for _ in range(args.iters + 3):
            flush.zero_()                                    # inputs (~180 MB at n=2) vs 126 MB L2: flush anyway
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[3:])
        return ts[len(ts) // 2]

    out = MSDeformAttnFunction.apply(value, sh, st, loc, attn, 64)
    fwd_ms = timed(lambda: MSDeformAttnFunction.apply(value.detach(), sh, st, loc.detach(), attn.detach(), 64))

    def bwd():
        o = MSDeformAttnFunction.apply(value, sh, st, loc, attn, 64)
        o.backward(go)
        value.grad = loc.grad = attn.grad = None

    fb_ms = timed(bwd)
    heads = n * lq * m
    fwd_bytes = heads * (l * p * (4 * d + 3) * 4 + d * 4)
    peak = 6544.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"op": "msda", "n": n, "queries": lq, "heads": m, "head_dim": d, "levels": l, "points": p,
                      "fwd_ms": fwd_ms, "fwd_bwd_ms": fb_ms, "fwd_algorithmic_gb": fwd_bytes / 1e9,
                      "fwd_gbs": fwd_bytes / fwd_ms / 1e6, "hbm_peak_gbs": peak, "fwd_frac": fwd_bytes / fwd_ms / 1e6 / peak,
                      "queries_per_s_fwd": n * lq / (fwd_ms / 1e3)}))


if __name__ == "__main__":
    main()
