"""oracle/msda_ref.py against golden vectors produced by the reference's own `ms_deform_attn_core_pytorch`
(functions/ms_deform_attn_func.py:41-61) and its autograd gradients — tests/golden/make_msda_golden.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_msda_golden import CASES, msda_case  # noqa: E402

from oracle import msda_ref  # noqa: E402

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "msda_golden.pt"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    value, shapes, starts, loc, attn, grad_out = msda_case(name)
    g = GOLD[name]
    out = msda_ref.msda_forward(value, shapes, starts, loc, attn)
    # the reference's own double-precision check uses torch.allclose defaults (ops/test.py:38)
    assert torch.allclose(out, g["output"]), (out - g["output"]).abs().max()
    gv, gl, ga = msda_ref.msda_backward(value, shapes, starts, loc, attn, grad_out)
    assert torch.allclose(gv, g["grad_value"], rtol=1e-9, atol=1e-12)
    assert torch.allclose(gl, g["grad_loc"], rtol=1e-9, atol=1e-12)
    assert torch.allclose(ga, g["grad_attn"], rtol=1e-9, atol=1e-12)


def test_oracle_float_matches_reference_tolerance():
    """ops/test.py:48-60: float32 forward within rtol 1e-2 / atol 1e-3 of the float64 reference."""
    value, shapes, starts, loc, attn, _ = msda_case("ops_test", torch.float32)
    out = msda_ref.msda_forward(value, shapes, starts, loc, attn)
    assert torch.allclose(out.double(), GOLD["ops_test"]["output"], rtol=1e-2, atol=1e-3)


def test_oracle_zero_padding_outside_the_map():
    """A sampling point more than one pixel outside contributes nothing (padding_mode='zeros')."""
    value = torch.ones(1, 6, 1, 1, dtype=torch.float64)
    loc = torch.tensor([-0.6, 0.5], dtype=torch.float64).view(1, 1, 1, 1, 1, 2)
    attn = torch.ones(1, 1, 1, 1, 1, dtype=torch.float64)
    assert float(msda_ref.msda_forward(value, [(2, 3)], [0], loc, attn)) == 0.0
    loc[..., 0] = 0.5
    assert float(msda_ref.msda_forward(value, [(2, 3)], [0], loc, attn)) == 1.0
