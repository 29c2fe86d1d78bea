"""Kernel-level parity of the tensor-core paths (conv_tc incl. halo mode, conv_tn, wgrad_tc incl. the 256-row and
tap-paired tiles, the bf16 stem, the fused optimizer kernels) against fp32 references: every case of
tools/gpu_diag.py, each in its own process (some cases select kernel variants through environment variables that the
library reads once)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["optim", "fwd_f32", "bwd_f32", "fwd_1x1_min", "fwd_1x1_k256", "fwd_1x1_n128", "fwd_3x3", "fwd_3x3_big",
         "fwd_epilogue", "fwd_epilogue2", "stride2_view", "fc", "stem", "bwd_1x1", "bwd_1x1_big", "bwd_3x3", "bwd_3x3_big",
         "bwd_mh2", "halo"]


@pytest.mark.parametrize("case", CASES)
def test_diag_case(case):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_diag.py"), "--case", case], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "FAIL" not in r.stdout, tail
