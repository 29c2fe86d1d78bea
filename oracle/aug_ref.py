"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's strong augmentation.

Only tests/ and bench/tools checker legs may import this; the product path is aldi_b200/augment.py -> csrc/augment.cu.

Follows aldi/aug.py:39-60 (`build_strong_augmentation`): on an HWC uint8 image, in this order
  1. RandomApply(p=0.8) of [RandomContrast(0.6,1.4), RandomBrightness(0.6,1.4), RandomSaturation(0.6,1.4)]
  2. RandomApply(p=0.2) of RandomSaturation(0, 0)                     ("random grayscale")
  3. RandomApply(p=0.5) of RandomBlurTransform((0.1, 2.0))            (aldi/aug.py:81-104, scipy gaussian_filter over
                                                                        ALL THREE axes of the HWC array, channels included)
  4. three RandomApply(p=0.7/0.5/0.3) of RandomEraseTransform(...)    (aldi/aug.py:106-152; labeled/unlabeled flags)
  5. optional MICTransform(ratio, block)                              (aldi/aug.py:154-186; AUG.*_MIC_AUG)
Steps 3-5 are the reference's own code and are PINNED by tests/golden/aug_golden.pt, produced by executing
aldi/aug.py's classes (tests/golden/make_aug_golden.py).  Steps 1-2 are Detectron2 transforms (absent here, "parity
unpinned"): restated from detectron2 v0.6 `RandomContrast/Brightness/Saturation` + fvcore `BlendTransform`
(uint8 -> float32, `src_weight*src_image + dst_weight*img`, clip, truncating cast back to uint8) with the NumPy-1.x
promotion D2 v0.6 ran under: contrast / brightness blend in float32, saturation in float64 (its `src_image` is a
float64 array).

The grayscale of RandomSaturation is `img.dot([0.299, 0.587, 0.114])`, which NumPy delegates to BLAS: its rounding
depends on the BLAS build / CPU / array shape (measured here: (N,3) and (H,W,3) views of the same pixels differ), so
that one step is reproducible only to 1 LSB after the truncating cast, in Detectron2 itself as much as here.

Random draws: `params_from_rngs` consumes Python `random` and `np.random` in exactly the order the reference's
transforms would (RandomApply and the D2 colour transforms draw from np.random; blur sigma and the erase rectangles
from Python `random`; erase fill and MIC mask from np.random), so the product's parameter drawing
(aldi_b200/augment.py) can be checked draw for draw.  The erase FILL itself (np.random.rand(h, w, c)) is replaced by
a counter-based hash on the device; parity there is on geometry and range, not on the noise values.
"""
import math
import random

import numpy as np
from scipy.ndimage import gaussian_filter


def blend(img_u8, src_image, src_weight, dst_weight, f64=False):
    """fvcore BlendTransform.apply_image for uint8 input."""
    img = img_u8.astype(np.float32)
    if f64:
        out = src_weight * src_image + (np.float32(dst_weight) * img).astype(np.float64)
    else:
        out = np.float32(np.float64(src_weight) * np.float64(src_image)) + np.float32(dst_weight) * img
    return np.clip(out, 0, 255).astype(np.uint8)


def color_jitter(img, cw, bw, sw):
    img = blend(img, img.mean(), 1 - cw, cw)                                  # RandomContrast
    img = blend(img, 0.0, 1 - bw, bw)                                          # RandomBrightness
    gray = img.dot([0.299, 0.587, 0.114])[:, :, np.newaxis]                    # RandomSaturation
    return blend(img, gray, 1 - sw, sw, f64=True)


def grayscale(img):
    gray = img.dot([0.299, 0.587, 0.114])[:, :, np.newaxis]
    return blend(img, gray, 1.0, 0.0, f64=True)


def blur(img, sigma):                                                          # aldi/aug.py:86-93
    out = gaussian_filter(img.astype(np.float32), sigma=sigma)
    return np.clip(out, 0, 255).astype(np.uint8)


def erase(img, rect, fill):                                                    # aldi/aug.py:118-143
    """rect = (h0, w0, h, w); fill: float array (h, w, c) in [0, 1) -> values fill*255, clipped, truncated."""
    h0, w0, h, w = rect
    out = img.astype(np.float32)
    out[h0:h0 + h, w0:w0 + w, :] = fill
    out[h0:h0 + h, w0:w0 + w, :] *= 255
    return np.clip(out, 0, 255).astype(np.uint8)


def mic_mask_to_pixels(mask, H, W):
    """cv2.resize(mask, (W, H), INTER_NEAREST): source index = min(floor(dst * ifx), src - 1) with
    ifx = 1 / (dst_size / src_size) in double (OpenCV resizeNN)."""
    mh, mw = mask.shape
    ys = np.minimum(np.floor(np.arange(H) * (1.0 / (H / mh))).astype(np.int64), mh - 1)
    xs = np.minimum(np.floor(np.arange(W) * (1.0 / (W / mw))).astype(np.int64), mw - 1)
    return mask[ys][:, xs]


def mic(img, mask):                                                            # aldi/aug.py:159-176
    H, W, _ = img.shape
    keep = mic_mask_to_pixels(mask.astype(np.uint8), H, W)
    return np.clip(img.astype(np.float32) * keep[..., None], 0, 255).astype(np.uint8)


ERASE_SPECS = ((0.7, 0.05, 0.2, 0.3, 3.3), (0.5, 0.02, 0.2, 0.1, 6.0), (0.3, 0.02, 0.2, 0.05, 8.0))   # aldi/aug.py:54-58


def draw_erase_rect(imgh, imgw, sl, sh, r1, r2):
    """The retry loop of RandomEraseTransform.apply_image (aldi/aug.py:124-137), Python `random` draws only.
    Returns (h0, w0, h, w) or None after 100 failed attempts."""
    for _ in range(100):
        area = imgw * imgh
        target_area = random.uniform(sl, sh) * area
        aspect_ratio = random.uniform(r1, r2)
        h = int(round(math.sqrt(target_area * aspect_ratio)))
        w = int(round(math.sqrt(target_area / aspect_ratio)))
        if w > 1 and h > 1 and w < imgw and h < imgh:
            h0 = random.randint(0, imgh - h - 1)
            w0 = random.randint(0, imgw - w - 1)
            return (h0, w0, h, w)
    return None


def params_from_rngs(H, W, include_erasing=True, mic=None, consume_fill=True):
    """Draw one image's augmentation parameters in the reference's RNG order.  mic = (ratio, block) or None."""
    p = {"color": None, "gray": False, "sigma": None, "erase": [], "mic": None}
    if np.random.uniform(0, 1.0) < 0.8:                                         # RandomApply._rand_range() < prob
        p["color"] = (np.random.uniform(0.6, 1.4), np.random.uniform(0.6, 1.4), np.random.uniform(0.6, 1.4))
    if np.random.uniform(0, 1.0) < 0.2:
        np.random.uniform(0, 0)                                                # RandomSaturation(0, 0) still draws
        p["gray"] = True
    if np.random.uniform(0, 1.0) < 0.5:
        p["sigma"] = random.uniform(0.1, 2.0)
    if include_erasing:
        for prob, sl, sh, r1, r2 in ERASE_SPECS:
            if np.random.uniform(0, 1.0) < prob:
                rect = draw_erase_rect(H, W, sl, sh, r1, r2)
                if rect is not None:
                    fill = np.random.rand(rect[2], rect[3], 3) if consume_fill else None
                    p["erase"].append((rect, fill))
    if mic is not None:
        np.random.uniform(0, 1.0)                                               # RandomApply(prob=1.0) still draws
        ratio, block = mic
        mh, mw = round(H / block), round(W / block)
        p["mic"] = np.random.rand(mh, mw) > ratio
    return p


def strong_augment(img_hwc_u8, p):
    img = img_hwc_u8
    if p["color"] is not None:
        img = color_jitter(img, *p["color"])
    if p["gray"]:
        img = grayscale(img)
    if p["sigma"] is not None:
        img = blur(img, p["sigma"])
    for rect, fill in p["erase"]:
        img = erase(img, rect, fill)
    if p["mic"] is not None:
        img = mic(img, p["mic"])
    return img
