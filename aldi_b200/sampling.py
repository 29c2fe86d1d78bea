"""Host-side definition of the device subsampler (csrc/select.cu: sample_hash / subsample_kernel /
roi_sample_kernel).

Reference behaviour (detectron2 modeling/sampling.py subsample_labels, driven by aldi/helpers.py:17-26
ManualSeed): `torch.randperm(n, device=model_device)[:k]` under the global generator.  CPU and CUDA randperm
streams differ, so the reference itself is not reproducible across devices.  The B200 path replaces
the permutation by a counter-based hash: candidate i gets the 32-bit key hash(seed, salt, i) and the k
candidates with the SMALLEST keys are taken (ties: lowest index), emitted in ascending index order.  Equal
(seed, salt, candidate set) => equal sample, which is exactly the property ManualSeed exists to provide for
the student/teacher RoI heads (aldi/distill.py:131-138).

This numpy model is what the parity tests install into the oracle so both sides draw the same samples.
"""
import numpy as np

SITE_RPN, SITE_ROI, SITE_RPN_DISTILL = 0, 1, 2


def make_salt(pass_id, site, image_idx):
    return ((int(pass_id) * 4 + int(site)) << 8 | int(image_idx)) & 0xFFFFFFFF


def _fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def sample_hash(seed, salt, index):
    """uint32 key of candidate `index` (array) — mirrors sample_hash() in csrc/select.cu."""
    with np.errstate(over="ignore"):
        s = _fmix32(np.array([(int(seed) ^ ((int(salt) * 0x27D4EB2F + 0x165667B1) & 0xFFFFFFFF)) & 0xFFFFFFFF],
                             dtype=np.uint32))[0]
        idx = np.asarray(index, dtype=np.uint64)
        v = ((idx * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ s
        return _fmix32(v)


def choose(seed, salt, is_negative, candidate_indices, take):
    """Positions (ascending) within `candidate_indices` of the `take` candidates the device would sample."""
    cand = np.asarray(candidate_indices, dtype=np.int64)
    if take <= 0 or cand.size == 0:
        return np.zeros(0, dtype=np.int64)
    keys = sample_hash(seed, (int(salt) * 2 + (1 if is_negative else 0)) & 0xFFFFFFFF, cand)
    order = np.lexsort((cand, keys))  # smallest key first, ties by lowest index
    return np.sort(order[:take])
