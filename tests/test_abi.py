"""The C-ABI library builds, loads on a CPU-only host and exports every symbol include/aldi_b200.h declares
(no compute calls here: those need a GPU and live in the `-m gpu` tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "aldi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aldi_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from aldi_b200 import build, lib
    build.build(verbose=False)
    return lib


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    assert len(syms) >= 30, syms
    for must in ("aldi_conv_tc", "aldi_wgrad_tc", "aldi_ema_update", "aldi_sgd_momentum_step", "aldi_nms_sorted",
                 "aldi_roi_align_forward", "aldi_distill_rpn_loss", "aldi_distill_roi_loss", "aldi_domain_bce_loss"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    cdll = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(cdll, s)]
    assert not missing, missing


def test_python_binding_covers_the_header(built_lib):
    declared = set(_declared_symbols())
    bound = set(built_lib.SIGNATURES)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    L = built_lib.load()
    assert L.aldi_abi_version() >= 1
    assert isinstance(L.aldi_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    monkeypatch.setattr(built_lib, "_lib", None)
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libaldi_b200.so")
    with pytest.raises(built_lib.AldiError):
        built_lib.load()
