"""Synthetic COCO-format batches for the benchmark and the parity tests (SURVEY.md §8d "Synthetic inputs").

Images are uint8 noise with planted solid-colour rectangles (the ground truth); the strong view applies a
deterministic colour jitter + MIC-style block masking with identical geometry, as aldi/aug.py:39-60,154-176
do on the dataloader workers (out of the hot path; this generator only has to produce the 4-tuple format
aldi/dataloader.py:57-80 yields).  Pure torch-CPU code so every platform regenerates identical bytes.
"""
import torch


def synth_image(h, w, gen, num_boxes=6, num_classes=8, min_side=16, max_side=None):
    max_side = max_side or max(min(h, w) // 2, min_side + 1)
    img = torch.randint(0, 256, (3, h, w), generator=gen, dtype=torch.int64).float()
    boxes, classes = [], []
    for _ in range(num_boxes):
        bw = int(torch.randint(min_side, max_side, (1,), generator=gen))
        bh = int(torch.randint(min_side, max_side, (1,), generator=gen))
        x0 = int(torch.randint(0, max(w - bw, 1), (1,), generator=gen))
        y0 = int(torch.randint(0, max(h - bh, 1), (1,), generator=gen))
        col = torch.randint(0, 256, (3, 1, 1), generator=gen).float()
        img[:, y0:y0 + bh, x0:x0 + bw] = col + torch.randn(3, bh, bw, generator=gen) * 4.0
        boxes.append([float(x0), float(y0), float(x0 + bw), float(y0 + bh)])
        classes.append(int(torch.randint(0, num_classes, (1,), generator=gen)))
    img = img.clamp(0, 255).round().to(torch.uint8)
    return img, torch.tensor(boxes, dtype=torch.float32), torch.tensor(classes, dtype=torch.int64)


def strong_view(img_u8, gen, mic_ratio=0.5, mic_block=32):
    """Colour jitter (brightness/contrast) + MIC block mask; geometry unchanged."""
    x = img_u8.float()
    b = 0.6 + 0.8 * float(torch.rand(1, generator=gen))
    c = 0.6 + 0.8 * float(torch.rand(1, generator=gen))
    mean = x.mean()
    x = ((x * b - mean) * c + mean).clamp(0, 255)
    _, h, w = x.shape
    mh, mw = (h + mic_block - 1) // mic_block, (w + mic_block - 1) // mic_block
    keep = (torch.rand(mh, mw, generator=gen) > mic_ratio).float()
    keep = keep.repeat_interleave(mic_block, 0).repeat_interleave(mic_block, 1)[:h, :w]
    return (x * keep).round().to(torch.uint8)


def synthetic_batch(seed, n_labeled, n_unlabeled, h, w, num_classes=8, num_boxes=6):
    """-> (labeled_strong, unlabeled_weak, unlabeled_strong): lists of dicts with plain tensors
    {"image": uint8 (3,h,w) BGR, "boxes": (G,4) xyxy fp32, "classes": (G,) int64, "height", "width"}."""
    gen = torch.Generator().manual_seed(seed)
    labeled, uw, us = [], [], []
    for _ in range(n_labeled):
        img, boxes, classes = synth_image(h, w, gen, num_boxes, num_classes)
        labeled.append({"image": strong_view(img, gen, mic_ratio=0.0), "boxes": boxes, "classes": classes,
                        "height": h, "width": w})
    for _ in range(n_unlabeled):
        img, boxes, classes = synth_image(h, w, gen, num_boxes, num_classes)
        uw.append({"image": img, "height": h, "width": w})
        us.append({"image": strong_view(img, gen), "height": h, "width": w})
    return labeled, uw, us
