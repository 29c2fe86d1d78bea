"""Golden vectors for the Deformable-DETR model + criterion from the REFERENCE's own code (run in the
authoring container only -- /root/reference does not exist on the GPU box).

The reference tree `aldi/detr/libs/DeformableDETRDetectron2/deformable_detr` is put on sys.path and imported as is:
`DeformableDETR` (models/deformable_detr.py:35-192) around `DeformableTransformer`
(models/deformable_transformer.py), `HungarianMatcher` (models/matcher.py), `SetCriterion`
(models/deformable_detr.py:195-370), `PositionEmbeddingSine`.  Two things are substituted, neither of them reference
arithmetic on this path:
  * the CUDA extension `MultiScaleDeformableAttention` (unbuildable here) is an empty stub and
    `MSDeformAttnFunction.apply` is routed to the reference's own pure-PyTorch `ms_deform_attn_core_pytorch`
    (functions/ms_deform_attn_func.py:41-61), which the reference's ops/test.py uses as the definition of the op;
  * the torchvision ResNet-50 trunk (pretrained download, no network) is replaced by a harness that hands fixed random
    multi-scale features + interpolated padding masks + the reference's sine position embeddings to the model in the
    (features, pos) format of `Joiner.forward` (models/backbone.py:112-129): everything downstream of the trunk -- input
    projections, the fourth level, encoder, decoder, heads, matching, losses, weights, inference top-k -- is the
    reference's code.
Weights: the reference's own initialisation followed by small noise on the tensors it zero-initialises (sampling
offsets' weight, attention weights, the last bbox layer), so no term of the forward or backward is identically zero.
Dropout 0 (it is a config value; with 0.1 the draws would be the only difference).

The trunk gets its own small golden (`detr_trunk_golden.pt`): the reference's `Backbone("resnet50", ...)` + `Joiner`
(models/backbone.py:25-129: torchvision ResNet-50 with the reference's FrozenBatchNorm2d, layer2-4 outputs, interpolated
masks, sine embeddings) on a 2 x 3 x 64 x 96 batch.  Only `pretrained=` is forced off (no network); the 23.5 M weights are
not stored but re-drawn from a seed by `trunk_state_dict()`, which the test shares.

    python tests/golden/make_detr_golden.py      ->  tests/golden/detr_golden.pt, tests/golden/detr_trunk_golden.pt
"""
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/aldi/detr/libs/DeformableDETRDetectron2"

CFG = dict(d_model=64, nhead=4, enc_layers=2, dec_layers=3, ffn=96, levels=4, points=2, queries=10, classes=5,
           channels=(16, 24, 32), strides=(8, 16, 32), canvas=(64, 96), sizes=((64, 96), (48, 80)),
           cost=(2.0, 5.0, 2.0), weights=dict(loss_ce=2.0, loss_bbox=5.0, loss_giou=2.0), focal_alpha=0.25, topk=7)


def detr_inputs(cfg=CFG, dtype=torch.float64):
    """features per level, the canvas padding mask, targets -- regenerated from the seed by the tests."""
    g = torch.Generator().manual_seed(5)
    H, W = cfg["canvas"]
    feats = [torch.randn(len(cfg["sizes"]), c, H // s, W // s, generator=g, dtype=torch.float64).to(dtype)
             for c, s in zip(cfg["channels"], cfg["strides"])]
    mask = torch.ones(len(cfg["sizes"]), H, W, dtype=torch.bool)
    for i, (h, w) in enumerate(cfg["sizes"]):
        mask[i, :h, :w] = False
    targets = []
    for n in (3, 2):
        cxcy = 0.2 + 0.6 * torch.rand(n, 2, generator=g, dtype=torch.float64)
        wh = 0.05 + 0.3 * torch.rand(n, 2, generator=g, dtype=torch.float64)
        targets.append({"labels": torch.randint(0, cfg["classes"], (n,), generator=g),
                        "boxes": torch.cat([cxcy, wh], 1).to(dtype)})
    return feats, mask, targets


def trunk_state_dict(shapes, dtype=torch.float64):
    """Seeded weights for every key of the reference trunk's state dict (sorted key order): convolutions ~ He-scaled
    normal, FrozenBN weight / var in [0.5, 1.5], bias / mean small."""
    g = torch.Generator().manual_seed(17)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("running_var") or (k.endswith("weight") and len(shp) == 1):
            v = 0.5 + torch.rand(shp, generator=g, dtype=torch.float64)
        elif len(shp) == 1:
            v = 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            v = torch.randn(shp, generator=g, dtype=torch.float64) * (2.0 / fan_in) ** 0.5
        sd[k] = v.to(dtype)
    return sd


def trunk_inputs(dtype=torch.float64):
    g = torch.Generator().manual_seed(23)
    x = torch.randn(2, 3, 64, 96, generator=g, dtype=torch.float64).to(dtype)
    mask = torch.ones(2, 64, 96, dtype=torch.bool)
    mask[0, :, :] = False
    mask[1, :40, :72] = False
    return x, mask


def load_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    import deformable_detr.models.deformable_detr as dd
    import deformable_detr.models.deformable_transformer as dt
    import deformable_detr.models.matcher as mt
    import deformable_detr.models.ops.functions.ms_deform_attn_func as fn
    import deformable_detr.models.ops.modules.ms_deform_attn as mod
    import deformable_detr.models.position_encoding as pe
    import deformable_detr.util.misc as misc

    class _Core:
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, im2col_step):
            return fn.ms_deform_attn_core_pytorch(value, shapes, loc, attn)

    mod.MSDeformAttnFunction = _Core
    return dd, dt, mt, pe, misc


class Harness(nn.Module):
    """Stands where `Joiner(Backbone, PositionEmbeddingSine)` does: (features, pos) from fixed feature maps."""

    def __init__(self, feats, pos_embed, misc, cfg):
        super().__init__()
        self.feats, self.pos_embed, self.misc = feats, pos_embed, misc
        self.strides, self.num_channels = list(cfg["strides"]), list(cfg["channels"])

    def __getitem__(self, i):
        assert i == 1
        return self.pos_embed

    def forward(self, samples):
        out, pos = [], []
        for x in self.feats:
            m = F.interpolate(samples.mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]    # models/backbone.py:91
            nt = self.misc.NestedTensor(x, m)
            out.append(nt)
            pos.append(self.pos_embed(nt).to(x.dtype))
        return out, pos


def main():
    cfg = CFG
    dd, dt, mt, pe, misc = load_reference()
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    feats, mask, targets = detr_inputs(cfg)
    transformer = dt.DeformableTransformer(d_model=cfg["d_model"], nhead=cfg["nhead"], num_encoder_layers=cfg["enc_layers"],
                                           num_decoder_layers=cfg["dec_layers"], dim_feedforward=cfg["ffn"], dropout=0.0,
                                           activation="relu", return_intermediate_dec=True, num_feature_levels=cfg["levels"],
                                           dec_n_points=cfg["points"], enc_n_points=cfg["points"], two_stage=False,
                                           two_stage_num_proposals=cfg["queries"])
    harness = Harness(feats, pe.PositionEmbeddingSine(cfg["d_model"] // 2, normalize=True), misc, cfg)
    model = dd.DeformableDETR(harness, transformer, num_classes=cfg["classes"], num_queries=cfg["queries"],
                              num_feature_levels=cfg["levels"], aux_loss=True, with_box_refine=False, two_stage=False)
    model = model.double()          # MSDeformAttn._reset_parameters builds its offset bias in float32 explicitly
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("sampling_offsets.weight", "attention_weights.weight", "attention_weights.bias")) or \
                    n.startswith("bbox_embed.0.layers.2"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    model.train()
    samples = misc.NestedTensor(torch.zeros(len(cfg["sizes"]), 3, *cfg["canvas"]), mask)
    out = model(samples)

    matcher = mt.HungarianMatcher(cost_class=cfg["cost"][0], cost_bbox=cfg["cost"][1], cost_giou=cfg["cost"][2])
    weight_dict = dict(cfg["weights"])                                    # meta_arch.py:109-120
    for i in range(cfg["dec_layers"] - 1):
        weight_dict.update({k + "_%d" % i: v for k, v in cfg["weights"].items()})
    weight_dict.update({k + "_enc": v for k, v in cfg["weights"].items()})
    criterion = dd.SetCriterion(cfg["classes"], matcher, weight_dict, ["labels", "boxes", "cardinality"],
                                focal_alpha=cfg["focal_alpha"])
    indices = [matcher({"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"]}, targets)]
    indices += [matcher(a, targets) for a in out["aux_outputs"]]
    losses = criterion(out, targets)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)       # meta_arch.py:161-167 + run_step's sum
    total.backward()

    gp = torch.Generator().manual_seed(21)
    grads = {}
    for n, p in model.named_parameters():
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        proj = torch.randn(p.shape, generator=gp)
        grads[n] = torch.stack([gr.norm(), (gr * proj).sum()])

    # inference (meta_arch.py:197-238): top-k over (query, class), boxes to absolute xyxy of each image
    prob = out["pred_logits"].detach().sigmoid()
    tv, ti = torch.topk(prob.view(prob.shape[0], -1), cfg["topk"], dim=1)

    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    golden = {
        "cfg": cfg, "state_dict": sd,
        "pred_logits": out["pred_logits"].detach(), "pred_boxes": out["pred_boxes"].detach(),
        "aux_logits": [a["pred_logits"].detach() for a in out["aux_outputs"]],
        "aux_boxes": [a["pred_boxes"].detach() for a in out["aux_outputs"]],
        "indices": [[(i.clone(), j.clone()) for i, j in lay] for lay in indices],
        "losses": {k: v.detach().clone() if torch.is_tensor(v) else torch.tensor(v) for k, v in losses.items()},
        "total": total.detach(), "grads": grads, "topk_scores": tv, "topk_index": ti,
    }
    path = os.path.join(HERE, "detr_golden.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(sd), "tensors;", {k: float(v) for k, v in golden["losses"].items()})


def main_trunk():
    import torchvision
    load_reference()
    import deformable_detr.models.backbone as bb
    import deformable_detr.models.position_encoding as pe
    import deformable_detr.util.misc as misc
    real = torchvision.models.resnet50

    def no_download(*a, **kw):
        kw["pretrained"] = False
        return real(*a, **kw)

    torchvision.models.resnet50 = no_download
    try:
        backbone = bb.Backbone("resnet50", True, True, False)
    finally:
        torchvision.models.resnet50 = real
    joiner = bb.Joiner(backbone, pe.PositionEmbeddingSine(128, normalize=True)).double().eval()
    shapes = {k: tuple(v.shape) for k, v in joiner.state_dict().items()}
    joiner.load_state_dict(trunk_state_dict(shapes), strict=True)
    x, mask = trunk_inputs()
    with torch.no_grad():
        feats, pos = joiner(misc.NestedTensor(x, mask))
    golden = {"shapes": shapes, "features": [f.tensors for f in feats], "masks": [f.mask for f in feats], "pos": pos,
              "trainable": sorted(n for n, p in joiner.named_parameters() if p.requires_grad)}
    path = os.path.join(HERE, "detr_trunk_golden.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(shapes), "tensors;", [tuple(f.shape) for f in golden["features"]])


if __name__ == "__main__":
    main()
    main_trunk()
