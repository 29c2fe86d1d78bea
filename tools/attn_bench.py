"""Attention kernels alone at the ViTDet-B shapes of BASELINE configs[2] (1024^2 input: 64 x 64 tokens, 12 heads): global
blocks (one 4096-token sequence per image) and windowed blocks (25 windows of 14 x 14 per image).  CUDA-event timing with an
L2 flush between iterations; `--ncu` runs each kernel ONCE (for `ncu --set full`, where no timing is printed).

    python tools/attn_bench.py [--images N] [--ncu] [--impl 0|1]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from aldi_b200 import lib as _l, ops  # noqa: E402


def problem(b, gh, gw, heads, dev):
    g = torch.Generator(device="cpu").manual_seed(0)
    t = gh * gw
    qkv = (torch.randn(b, t, 3 * heads * 64, generator=g) * 0.5).bfloat16().to(dev)
    rel_h = (torch.randn(b, heads, gh, t, generator=g) * 0.2).to(dev)
    rel_w = (torch.randn(b, heads, gw, t, generator=g) * 0.2).to(dev)
    dout = torch.randn(b, t, heads * 64, generator=g).bfloat16().to(dev)
    out = torch.zeros(b, t, heads * 64, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(b, heads, t, device=dev)
    dqkv, drel_h, drel_w, delta = torch.zeros_like(qkv), torch.zeros_like(rel_h), torch.zeros_like(rel_w), torch.zeros_like(lse)
    p = _l.AttnParams()
    p.qkv, p.batch, p.gh, p.gw, p.heads = qkv.data_ptr(), b, gh, gw, heads
    p.row_stride, p.batch_stride = qkv.stride(1), qkv.stride(0)
    p.rel_h, p.rel_w, p.scale, p.dtype = rel_h.data_ptr(), rel_w.data_ptr(), 0.125, _l.BF16
    p.out, p.out_stride, p.out_batch_stride, p.lse = out.data_ptr(), out.stride(1), out.stride(0), lse.data_ptr()
    p.dout, p.dqkv, p.drel_h, p.drel_w, p.delta = dout.data_ptr(), dqkv.data_ptr(), drel_h.data_ptr(), drel_w.data_ptr(), delta.data_ptr()
    keep = (qkv, rel_h, rel_w, dout, out, lse, dqkv, drel_h, drel_w, delta)
    return p, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1)
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _l.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    for name, (b, gh, gw) in (("global 64x64", (args.images, 64, 64)), ("window 14x14", (25 * args.images, 14, 14))):
        p, keep = problem(b, gh, gw, 12, dev)
        p.impl = args.impl
        t = gh * gw
        fl_f, fl_b = 4.0 * b * 12 * t * t * 64, 10.0 * b * 12 * t * t * 64
        for what, fn, fl in (("forward", L.aldi_attention_forward, fl_f), ("backward", L.aldi_attention_backward, fl_b)):
            _l.check(fn(ctypes.byref(p), ops._stream()), what)          # warm-up (and the only launch under ncu)
            torch.cuda.synchronize()
            if args.ncu:
                continue
            ms = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _l.check(fn(ctypes.byref(p), ops._stream()), what)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            ms.sort()
            med = ms[len(ms) // 2]
            print(json.dumps({"shape": name, "batch": b, "kernel": what, "impl": args.impl, "ms": round(med, 4),
                              "tflops": round(fl / med / 1e9, 1), "frac_of_bf16_sustained_peak": round(fl / med / 1e9 / peak, 4)}),
                  flush=True)


if __name__ == "__main__":
    main()
