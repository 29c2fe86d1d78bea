"""oracle/convnext_ref.py against goldens from the reference's own ConvNeXt class (tests/golden/make_convnext_golden.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_convnext_golden import CASES, MEAN, STD, convnext_case, grad_summary  # noqa: E402

from oracle import convnext_ref  # noqa: E402

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "convnext_golden.pt"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_convnext_matches_reference(name):
    seed, depths, dims, dpr, ls, n, h, w = CASES[name]
    sd, images, keeps, gouts = convnext_case(name)
    m = convnext_ref.ConvNeXt(depths=depths, dims=dims, drop_path_rate=dpr, layer_scale_init_value=ls).double()
    m.load_state_dict(sd, strict=True)
    m.train()
    m.keep_queue = [k for k in keeps if k is not None]
    x = (images.double() - torch.tensor(MEAN, dtype=torch.float64).view(1, 3, 1, 1)) / torch.tensor(STD, dtype=torch.float64).view(1, 3, 1, 1)
    outs = m(x)
    assert not m.keep_queue
    g = GOLD[name]
    for i in range(4):
        assert torch.allclose(outs[i].float(), g["outs"][i], rtol=1e-5, atol=1e-6)
    sum((outs[i] * gouts[i]).sum() for i in range(4)).backward()
    for k, p in m.named_parameters():
        if "grads" in g:
            assert torch.allclose(p.grad.float(), g["grads"][k], rtol=1e-4, atol=1e-6), k
        else:
            norm, proj = g["grad_summary"][k]
            mine = grad_summary(k, p.grad)
            assert abs(mine[0] - norm) <= 1e-6 * norm + 1e-12 and abs(mine[1] - proj) <= 1e-6 * norm * p.numel() ** 0.5 + 1e-12, k
