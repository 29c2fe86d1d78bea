"""Strong augmentation on the device: host half.

The reference builds the strong view of every image on dataloader workers (aldi/aug.py:16-60 `get_augs` /
`build_strong_augmentation`, NumPy + scipy + cv2) and ships BOTH views over PCIe.  Here the weak view is the only
image that crosses the bus; `StrongAugmenter.apply` derives the strong one in HBM with csrc/augment.cu.  The random
DECISIONS stay on the host and are drawn in the reference's order from the reference's generators (`np.random` for
RandomApply and the Detectron2 colour transforms, Python `random` for the blur sigma and the erase rectangles,
`np.random` for the MIC mask), so a run seeded like the reference takes the same decisions; only the erase fill
noise comes from a device-side hash instead of `np.random.rand(h, w, c)` (which the reference still consumes from the
stream — `draw` can skip or burn those draws).
"""
import ctypes
import math
import random
import zlib

import numpy as np
import torch

from . import lib as _l

ERASE_SPECS = ((0.7, 0.05, 0.2, 0.3, 3.3), (0.5, 0.02, 0.2, 0.1, 6.0), (0.3, 0.02, 0.2, 0.05, 8.0))   # aldi/aug.py:54-58


def gaussian_taps(sigma):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(4 * sigma + 0.5)."""
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return radius, phi / phi.sum()


class StrongAugmenter:
    def __init__(self, labeled, include_erasing=True, mic=None, burn_fill_draws=True):
        """labeled / include_erasing / mic mirror get_augs(cfg, labeled, ...) (aldi/aug.py:16-37): `include_erasing` is
        AUG.{LABELED,UNLABELED}_INCLUDE_RANDOM_ERASING, mic = (AUG.MIC_RATIO, AUG.MIC_BLOCK_SIZE) when the MIC flag of
        that split is on.  burn_fill_draws keeps np.random in step with the reference (it draws h*w*c fill values)."""
        self.labeled, self.include_erasing, self.mic, self.burn = labeled, include_erasing, mic, burn_fill_draws
        self._ws = {}
        self._fills = 0

    @classmethod
    def from_config(cls, cfg, labeled):
        a = cfg.AUG
        erasing = a.LABELED_INCLUDE_RANDOM_ERASING if labeled else a.UNLABELED_INCLUDE_RANDOM_ERASING
        mic_on = a.LABELED_MIC_AUG if labeled else a.UNLABELED_MIC_AUG
        return cls(labeled, erasing, (a.MIC_RATIO, a.MIC_BLOCK_SIZE) if mic_on else None)

    # ---- parameters: same draws, same order as the reference's transform list --------------------------------------
    def draw(self, h, w):
        p = {"color": None, "gray": False, "sigma": None, "erase": [], "mic": None}
        if np.random.uniform(0, 1.0) < 0.8:
            p["color"] = (np.random.uniform(0.6, 1.4), np.random.uniform(0.6, 1.4), np.random.uniform(0.6, 1.4))
        if np.random.uniform(0, 1.0) < 0.2:
            np.random.uniform(0, 0)
            p["gray"] = True
        if np.random.uniform(0, 1.0) < 0.5:
            p["sigma"] = random.uniform(0.1, 2.0)
        if self.include_erasing:
            for prob, sl, sh, r1, r2 in ERASE_SPECS:
                if np.random.uniform(0, 1.0) < prob:
                    rect = self._erase_rect(h, w, sl, sh, r1, r2)
                    if rect is not None:
                        if self.burn:
                            np.random.rand(rect[2], rect[3], 3)
                        # fill-noise seed: derived, so that neither reference RNG stream advances
                        self._fills += 1
                        p["erase"].append((rect, zlib.crc32(repr((rect, self._fills)).encode())))
        if self.mic is not None:
            np.random.uniform(0, 1.0)
            ratio, block = self.mic
            p["mic"] = np.random.rand(round(h / block), round(w / block)) > ratio
        return p

    @staticmethod
    def _erase_rect(imgh, imgw, sl, sh, r1, r2):
        for _ in range(100):                                    # aldi/aug.py:124-137
            target_area = random.uniform(sl, sh) * (imgw * imgh)
            aspect_ratio = random.uniform(r1, r2)
            h = int(round(math.sqrt(target_area * aspect_ratio)))
            w = int(round(math.sqrt(target_area / aspect_ratio)))
            if w > 1 and h > 1 and w < imgw and h < imgh:
                return (random.randint(0, imgh - h - 1), random.randint(0, imgw - w - 1), h, w)
        return None

    # ---- device --------------------------------------------------------------------------------------------------
    def apply(self, src, dst, params, valid_hw=None):
        """src, dst: uint8 (3, H, W) device tensors (views into a padded canvas are fine); params from `draw`.
        valid_hw: the image's own (h, w) inside the canvas (default: the whole tensor)."""
        if not (src.is_cuda and dst.is_cuda):
            raise _l.AldiError("StrongAugmenter.apply runs on the GPU only (there is no CPU fallback)")
        assert src.dtype == torch.uint8 and dst.dtype == torch.uint8 and src.dim() == 3 and src.shape[0] == 3
        assert src.stride(2) == 1 and dst.stride(2) == 1
        h, w = valid_hw if valid_hw is not None else (src.shape[1], src.shape[2])
        q = _l.AugParams()
        q.h, q.w = int(h), int(w)
        q.src_plane, q.src_row, q.dst_plane, q.dst_row = src.stride(0), src.stride(1), dst.stride(0), dst.stride(1)
        q.do_color = int(params["color"] is not None)
        if params["color"] is not None:
            q.contrast_w, q.brightness_w, q.saturation_w = (float(v) for v in params["color"])
        q.do_gray = int(bool(params["gray"]))
        q.blur_radius = -1
        if params["sigma"] is not None:
            radius, taps = gaussian_taps(params["sigma"])
            q.blur_radius = radius
            for i, t in enumerate(taps):
                q.blur_taps[i] = float(t)
            assert src.data_ptr() != dst.data_ptr(), "blur needs distinct source and destination"
        q.num_erase = len(params["erase"])
        for e, (rect, seed) in enumerate(params["erase"]):
            for k in range(4):
                q.erase_rect[e][k] = int(rect[k])
            q.erase_seed[e] = int(seed) & 0xFFFFFFFF
        keep = None
        if params["mic"] is not None:
            keep = torch.from_numpy(np.ascontiguousarray(params["mic"]).astype(np.uint8)).to(src.device, non_blocking=True)
            q.mic_mask, q.mic_h, q.mic_w = keep.data_ptr(), int(keep.shape[0]), int(keep.shape[1])
        L = _l.load()
        nbytes = int(L.aldi_strong_augment_workspace_bytes(q.h, q.w))
        # one workspace per (device, size, STREAM): apply() is called from the compute stream and from the copy stream
        # (prefetch of the next micro-batch), and two streams must not share the blur scratch planes
        cur = torch.cuda.current_stream()
        key = (src.device, nbytes, cur.cuda_stream)
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
        stream = ctypes.c_void_p(cur.cuda_stream)
        _l.check(L.aldi_strong_augment(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), ctypes.byref(q),
                                       ctypes.c_void_p(ws.data_ptr()), nbytes, stream), "aldi_strong_augment")
        return dst
