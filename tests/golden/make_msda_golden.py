"""Golden vectors for multi-scale deformable attention from the REFERENCE's own pure-PyTorch implementation
(run in the authoring container only; /root/reference does not exist on the GPU box).

`functions/ms_deform_attn_func.py` is loaded from the reference tree with its `import MultiScaleDeformableAttention`
(the CUDA extension, unbuildable here: SURVEY §8c) replaced by an empty stub; `ms_deform_attn_core_pytorch`
(:41-61) then runs in float64 and torch.autograd provides the reference gradients.  Inputs are regenerated from the
seeds by `msda_case()` below, so only outputs / gradients are stored.

    python tests/golden/make_msda_golden.py      ->  tests/golden/msda_golden.pt
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/aldi/detr/libs/DeformableDETRDetectron2/deformable_detr/models/ops/functions/ms_deform_attn_func.py"

# name -> (seed, N, M, D, Lq, P, shapes, location range)
CASES = {
    # the reference's own known-answer shapes and seed (ops/test.py:21-28)
    "ops_test": (3, 1, 2, 2, 2, 2, ((6, 4), (3, 2)), (0.0, 1.0)),
    # gradcheck channel counts of ops/test.py:86 that are cheap on CPU
    "ops_test_c30": (3, 1, 2, 30, 2, 2, ((6, 4), (3, 2)), (0.0, 1.0)),
    "ops_test_c71": (3, 1, 2, 71, 2, 2, ((6, 4), (3, 2)), (0.0, 1.0)),
    # Deformable-DETR geometry (d_model 256, 8 heads, 4 levels, 4 points), small maps; locations reach outside [0, 1]
    # so the zero-padding branches of every tap are exercised
    "detr_small": (11, 1, 8, 32, 37, 4, ((12, 16), (6, 8), (3, 4), (2, 2)), (-0.15, 1.15)),
}


def msda_case(name, dtype=torch.float64):
    seed, n, m, d, lq, p, shapes, (lo, hi) = CASES[name]
    g = torch.Generator().manual_seed(seed)
    s = sum(h * w for h, w in shapes)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    value = (torch.rand(n, s, m, d, generator=g, dtype=torch.float64) * 0.01).to(dtype)
    loc = (lo + (hi - lo) * torch.rand(n, lq, m, len(shapes), p, 2, generator=g, dtype=torch.float64)).to(dtype)
    attn = torch.rand(n, lq, m, len(shapes), p, generator=g, dtype=torch.float64) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dtype)
    grad_out = torch.randn(n, lq, m * d, generator=g, dtype=torch.float64).to(dtype)
    return value, list(shapes), starts, loc, attn, grad_out


def load_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("ref_ms_deform_attn_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def main():
    core = load_reference()
    out = {}
    for name in CASES:
        value, shapes, starts, loc, attn, grad_out = msda_case(name)
        value.requires_grad_(True)
        loc.requires_grad_(True)
        attn.requires_grad_(True)
        y = core(value, torch.as_tensor(shapes, dtype=torch.long), loc, attn)
        y.backward(grad_out)
        out[name] = {"output": y.detach().clone(), "grad_value": value.grad.clone(), "grad_loc": loc.grad.clone(),
                     "grad_attn": attn.grad.clone()}
        print(name, tuple(y.shape), float(y.abs().max()))
    torch.save(out, os.path.join(HERE, "msda_golden.pt"))


if __name__ == "__main__":
    main()
