"""Independent cross-checks of the Detectron2 restatement (oracle/d2_rcnn.py, parity UNPINNED: detectron2 is not under
/root/reference) against torchvision's detection utilities -- written by other hands, same maskrcnn-benchmark lineage, shipped
in this image: anchor matching incl. low-quality matches, the box coder, FrozenBatchNorm, the RoI level mapper, the multi-level
RoIAlign pooler.  This does not pin the oracle to the reference (nothing can here); it shows the restated pieces agree with a
second published implementation wherever the two define the same function."""
import math

import pytest
import torch

from oracle import d2_rcnn as d2

tvu = pytest.importorskip("torchvision.models.detection._utils")


def _boxes(n, g, size=400.0):
    xy = torch.rand(n, 2, generator=g) * size
    wh = torch.rand(n, 2, generator=g) * size / 2 + 1.0
    return torch.cat([xy, xy + wh], dim=1)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matcher_with_low_quality_matches_agrees_with_torchvision(seed):
    g = torch.Generator().manual_seed(seed)
    gt, anchors = _boxes(7, g), _boxes(500, g)
    iou = d2.pairwise_iou(d2.Boxes(gt), d2.Boxes(anchors))
    matches, labels = d2.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(iou)
    tv = tvu.Matcher(0.7, 0.3, allow_low_quality_matches=True)(iou)
    want = torch.where(tv >= 0, torch.ones_like(tv), torch.where(tv == tvu.Matcher.BELOW_LOW_THRESHOLD, torch.zeros_like(tv),
                                                                 -torch.ones_like(tv)))
    assert torch.equal(labels.long(), want)
    assert torch.equal(matches[tv >= 0], tv[tv >= 0])
    # the RoI-head matcher: one threshold, no low-quality matches
    m2, l2 = d2.Matcher([0.5], [0, 1], allow_low_quality_matches=False)(iou)
    tv2 = tvu.Matcher(0.5, 0.5, allow_low_quality_matches=False)(iou)
    assert torch.equal(l2.long(), (tv2 >= 0).long()) and torch.equal(m2[tv2 >= 0], tv2[tv2 >= 0])


@pytest.mark.parametrize("weights", [(1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)])
def test_box2box_transform_agrees_with_torchvision_box_coder(weights):
    g = torch.Generator().manual_seed(3)
    src, tgt = _boxes(300, g), _boxes(300, g)
    clamp = math.log(1000.0 / 16)
    ours = d2.Box2BoxTransform(weights, scale_clamp=clamp)
    tv = tvu.BoxCoder(weights, bbox_xform_clip=clamp)
    deltas = ours.get_deltas(src, tgt)
    assert torch.allclose(deltas, tv.encode_single(tgt, src), rtol=1e-6, atol=1e-6)
    big = deltas * 3.0                                            # some log-scales beyond the clamp
    assert torch.allclose(ours.apply_deltas(big, src), tv.decode_single(big, src), rtol=1e-5, atol=1e-3)
    ok = (deltas[:, 2] / weights[2] < clamp) & (deltas[:, 3] / weights[3] < clamp)       # round trip where the clamp is idle
    assert ok.sum() > 200 and torch.allclose(ours.apply_deltas(deltas, src)[ok], tgt[ok], rtol=1e-4, atol=1e-2)


def test_frozen_batchnorm_agrees_with_torchvision():
    from torchvision.ops.misc import FrozenBatchNorm2d
    g = torch.Generator().manual_seed(4)
    a, b = d2.FrozenBatchNorm2d(16), FrozenBatchNorm2d(16, eps=1e-5)
    sd = {"weight": torch.randn(16, generator=g), "bias": torch.randn(16, generator=g), "running_mean": torch.randn(16, generator=g),
          "running_var": torch.rand(16, generator=g) + 0.1}
    a.load_state_dict(sd, strict=False)
    b.load_state_dict(sd, strict=False)
    x = torch.randn(2, 16, 5, 7, generator=g)
    assert torch.allclose(a(x), b(x), rtol=1e-6, atol=1e-6)


def test_roi_level_assignment_and_pooler_agree_with_torchvision():
    from torchvision.ops import MultiScaleRoIAlign
    from torchvision.ops.poolers import LevelMapper
    g = torch.Generator().manual_seed(5)
    boxes = [_boxes(200, g, 600.0), _boxes(150, g, 600.0)]
    lv = d2.assign_boxes_to_levels([d2.Boxes(b) for b in boxes], 2, 5, 224, 4)
    tv = LevelMapper(2, 5, canonical_scale=224, canonical_level=4)(boxes)
    # the two put their epsilon in different places; away from the level boundaries they are the same map
    s = torch.sqrt(torch.cat([(b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) for b in boxes]))
    frac = (4 + torch.log2(s / 224)) % 1.0
    safe = (frac > 1e-4) & (frac < 1 - 1e-4)
    assert torch.equal(lv[safe], tv[safe])
    feats = {"p%d" % l: torch.randn(2, 8, 640 // 2 ** l, 640 // 2 ** l, generator=g) for l in (2, 3, 4, 5)}
    ours = d2.ROIPooler(7, (1 / 4, 1 / 8, 1 / 16, 1 / 32), 0, 224, 4)([feats["p%d" % l] for l in (2, 3, 4, 5)],
                                                                      [d2.Boxes(b) for b in boxes])
    # torchvision's pooler is RoIAlign v1 (aligned=False): compare its LEVEL ROUTING through per-level aligned calls instead
    import torchvision
    fmt = d2.convert_boxes_to_pooler_format([d2.Boxes(b) for b in boxes])
    want = torch.zeros_like(ours)
    for i, l in enumerate((2, 3, 4, 5)):
        idx = (tv == i).nonzero().flatten()
        want[idx] = torchvision.ops.roi_align(feats["p%d" % l], fmt[idx], (7, 7), 1.0 / 2 ** l, 0, aligned=True)
    assert torch.allclose(ours[safe], want[safe], rtol=1e-6, atol=1e-6)
    assert isinstance(MultiScaleRoIAlign(["p2"], 7, 0), torch.nn.Module)      # the class whose mapper was used above
