"""oracle/detr_ref.py against tests/golden/detr_golden.pt, which tests/golden/make_detr_golden.py produced by executing
the reference's own DeformableDETR / DeformableTransformer / HungarianMatcher / SetCriterion classes in float64."""
import os
import sys

import torch

from oracle import detr_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_detr_golden import detr_inputs, trunk_inputs, trunk_state_dict  # noqa: E402


def _load():
    g = torch.load(os.path.join(HERE, "golden", "detr_golden.pt"), weights_only=False)
    feats, mask, targets = detr_inputs(g["cfg"])
    return g, feats, mask, targets


def test_forward_every_decoder_layer():
    g, feats, mask, targets = _load()
    with torch.no_grad():
        out = R.forward(g["state_dict"], g["cfg"], feats, mask)
    assert torch.allclose(out["pred_logits"], g["pred_logits"], atol=1e-9, rtol=1e-9)
    assert torch.allclose(out["pred_boxes"], g["pred_boxes"], atol=1e-9, rtol=1e-9)
    for a, gl, gb in zip(out["aux_outputs"], g["aux_logits"], g["aux_boxes"]):
        assert torch.allclose(a["pred_logits"], gl, atol=1e-9, rtol=1e-9) and torch.allclose(a["pred_boxes"], gb, atol=1e-9, rtol=1e-9)


def test_matching_losses_gradients_and_topk():
    g, feats, mask, targets = _load()
    cfg = g["cfg"]
    # the shared heads appear three times in the state dict; only index 0 is read and receives the gradient
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["state_dict"].items()}
    out = R.forward(sd, cfg, feats, mask)
    losses, ind = R.criterion(out, targets, cfg)
    # Hungarian assignment of the final and of each auxiliary layer: exact
    assert len(ind) == len(g["indices"]) == cfg["dec_layers"]
    for mine, ref in zip(ind, g["indices"]):
        for (i, j), (ri, rj) in zip(mine, ref):
            assert torch.equal(i, ri) and torch.equal(j, rj)
    # every entry of the reference's loss dict, incl. the unweighted logging entries
    assert set(losses) == set(g["losses"])
    for k, v in g["losses"].items():
        assert abs(float(losses[k].detach()) - float(v)) <= 1e-9 * max(1.0, abs(float(v))), k
    w = R.weight_dict(cfg)
    assert {"loss_ce_enc", "loss_bbox_1", "loss_giou_0"} <= set(w) and "class_error" not in w
    total = sum(losses[k] * w[k] for k in losses if k in w)
    assert abs(float(total.detach()) - float(g["total"])) <= 1e-9 * abs(float(g["total"]))
    total.backward()
    gp = torch.Generator().manual_seed(21)
    for name, ref in g["grads"].items():                  # named_parameters order of the reference model
        p = sd[name]
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        proj = torch.randn(p.shape, generator=gp, dtype=torch.float64)
        got = torch.stack([gr.norm(), (gr * proj).sum()])
        assert torch.allclose(got, ref, atol=1e-9, rtol=1e-7), (name, got, ref)
    assert float(g["grads"]["transformer.encoder.layers.0.self_attn.sampling_offsets.weight"][0]) > 0
    scores, idx, labels, boxes = R.inference_topk(out["pred_logits"].detach(), out["pred_boxes"].detach(), cfg["sizes"], cfg["topk"])
    assert torch.equal(idx, g["topk_index"]) and torch.allclose(scores, g["topk_scores"], atol=1e-12)
    assert boxes.shape == (2, cfg["topk"], 4) and int(labels.max()) < cfg["classes"]


def test_position_embedding_ignores_padding_and_empty_targets():
    mask = torch.ones(1, 4, 6, dtype=torch.bool)
    mask[0, :3, :5] = False
    pos = R.sine_position_embedding(mask, 8)
    assert pos.shape == (1, 16, 4, 6)
    # the normaliser is the LAST cumulative coordinate: the valid extent (3 rows, 5 columns), not the canvas
    full = R.sine_position_embedding(torch.zeros(1, 3, 5, dtype=torch.bool), 8)
    assert torch.allclose(pos[:, :, :3, :5], full, atol=1e-6)
    # an image without ground truth: num_boxes clamps to 1, box losses are 0, the class loss is all-negative focal
    g, feats, mask2, targets = _load()
    with torch.no_grad():
        out = R.forward(g["state_dict"], g["cfg"], feats, mask2)
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros(0, 4, dtype=torch.float64)} for _ in range(2)]
    losses, ind = R.criterion(out, empty, g["cfg"])
    assert all(len(i) == 0 for lay in ind for i, _ in lay)
    assert float(losses["loss_bbox"]) == 0.0 and float(losses["loss_giou_1"]) == 0.0 and float(losses["loss_ce"]) > 0


def test_trunk_matches_reference_backbone_and_joiner():
    g = torch.load(os.path.join(HERE, "golden", "detr_trunk_golden.pt"), weights_only=False)
    assert R.trunk_shapes() == g["shapes"]                      # every key and shape of the reference trunk's state dict
    assert R.trunk_trainable() == g["trainable"]                # layer2-4 convolutions only
    sd = trunk_state_dict(R.trunk_shapes())
    x, mask = trunk_inputs()
    with torch.no_grad():
        feats, masks, pos = R.trunk(sd, x, mask)
    assert [tuple(f.shape) for f in feats] == [(2, 512, 8, 12), (2, 1024, 4, 6), (2, 2048, 2, 3)]
    for f, m, p, gf, gm, gpos in zip(feats, masks, pos, g["features"], g["masks"], g["pos"]):
        assert torch.allclose(f, gf, rtol=1e-9, atol=1e-9 * float(gf.abs().max()))
        assert torch.equal(m, gm) and torch.allclose(p, gpos, atol=1e-12)
    assert bool(masks[0][1, 5:, :].all()) and not bool(masks[0][1, :5, :9].any())      # 40 x 72 valid pixels of 64 x 96
