"""aldi_b200 — the ALDI++ teacher-student training step on NVIDIA B200 (sm_100a).

Host code in Python over PyTorch tensors (memory / streams / torch.distributed only); all arithmetic of the
per-iteration hot path runs in hand-written CUDA kernels behind the C ABI in include/aldi_b200.h
(libaldi_b200.so, built in-tree by `python -m aldi_b200.build`).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
