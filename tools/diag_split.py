"""GPU diagnostic: the split-bf16 tensor-core mode against the CUDA-core fp32 mode, layer by layer (first mismatch wins)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import parity_utils as pu  # noqa: E402
from aldi_b200 import ops  # noqa: E402
from aldi_b200.detector import Detector, DetectorWeights, FlatLayout  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    dev = torch.device("cuda")
    # single conv calls first
    g = torch.Generator().manual_seed(0)
    for (n, h, w, cin, cout, k) in ((1, 16, 24, 64, 64, 1), (2, 20, 24, 64, 64, 3), (1, 16, 16, 256, 256, 3), (1, 8, 8, 128, 512, 1)):
        x = torch.randn(n, h, w, cin, generator=g).to(dev)
        wt = (torch.randn(cout, k * k * cin, generator=g) / (k * k * cin) ** 0.5).to(dev)
        ref = torch.empty(n, h, w, cout, device=dev)
        ops.conv(x, wt, ref, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2, relu=True)
        for parts in (2, 3):
            out = torch.empty_like(ref)
            ops.conv(x, ops.split_bf16(wt, parts), out, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2, relu=True)
            print("conv %dx%d %d->%d parts %d: rel err %.3e" % (k, k, cin, cout, parts, rel(out, ref)), flush=True)
    sd_s, sd_t, ls, uw, us = pu.make_inputs(7, 1, 1, 128, 160)
    layout = FlatLayout(8)
    flat = layout.pack_state_dict(sd_s).to(dev)
    det = Detector(8)
    img = uw[0]["image"].unsqueeze(0).to(dev)
    sizes = torch.tensor([[128, 160]], dtype=torch.int32, device=dev)
    outs = {}
    for tag, parts in (("fp32", 0), ("x6", 3), ("x3", 2)):
        W = DetectorWeights(layout, flat, torch.float32, split_parts=parts)
        W.refresh()
        feats, _ = det.backbone(W, img, sizes, save=False)
        lv = det.levels(feats)
        rpn_out, _ = det.rpn_head(W, feats, lv, save=False)
        outs[tag] = dict(feats, rpn_out=rpn_out)
    for tag in ("x6", "x3"):
        for k in ("res2", "res3", "res4", "res5", "p5", "p4", "p3", "p2", "rpn_out"):
            print("%s %-8s rel err vs fp32 mode: %.3e" % (tag, k, rel(outs[tag][k], outs["fp32"][k])), flush=True)


if __name__ == "__main__":
    main()
