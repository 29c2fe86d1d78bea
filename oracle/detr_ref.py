"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Deformable-DETR path (BASELINE configs[3], SURVEY §8 f4 /
Appendix B): torchvision-style ResNet-50 trunk with FrozenBN + padding masks + sine position embedding, input
projections + fourth level, deformable encoder / decoder, heads, Hungarian matching, set criterion, inference top-k.

Only tests/ may import this module.  Reference files (under
/root/reference/aldi/detr/libs/DeformableDETRDetectron2/, "DETR/" below) each function follows are cited inline.
PINNED: tests/golden/make_detr_golden.py executes the reference's own `DeformableDETR`, `DeformableTransformer`,
`HungarianMatcher` and `SetCriterion` classes in float64 (the CUDA op routed to the reference's pure-PyTorch core, the
torchvision trunk replaced by fixed feature maps) and stores outputs of every decoder layer, the assignment of every
layer, all loss entries, the weighted total, a norm + projection of every parameter gradient and the inference top-k;
tests/test_detr_oracle.py replays them against this file.  The trunk has its own golden from the reference's
`Backbone` + `Joiner` classes (`detr_trunk_golden.pt`; weights re-drawn from a seed on both sides).

Functional on purpose: the model is a plain dict of tensors under the reference's state-dict keys
(`transformer.encoder.layers.0.self_attn.sampling_offsets.weight`, `input_proj.3.0.weight`, `class_embed.0.bias`, ...),
which is the form a flat-buffer CUDA implementation consumes.  Without box refinement the reference shares ONE class /
box head between all decoder layers (DETR/deformable_detr/models/deformable_detr.py:105-107): the keys `class_embed.{i}` /
`bbox_embed.{i}` repeat the same tensors and only index 0 is read here.
"""
import math

import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from .msda_ref import msda_forward


# --------------------------------------------------------------------------------------------------------------
def sine_position_embedding(mask, num_pos_feats, temperature=10000.0, scale=2 * math.pi):
    """DETR/deformable_detr/models/position_encoding.py:36-56 (normalize=True): cumulative coordinates over the
    NON-padded pixels in float32, (coord - 0.5) / (last + 1e-6) * 2 pi, interleaved sin / cos, cat(pos_y, pos_x)."""
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    y = (y - 0.5) / (y[:, -1:, :] + 1e-6) * scale
    x = (x - 0.5) / (x[:, :, -1:] + 1e-6) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / num_pos_feats)

    def enc(c):
        p = c[:, :, :, None] / dim_t
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).flatten(3)

    return torch.cat((enc(y), enc(x)), dim=3).permute(0, 3, 1, 2)


def level_mask(canvas_mask, hw):
    """nearest-neighbour resize of the canvas padding mask (DETR/deformable_detr/models/backbone.py:91)."""
    return F.interpolate(canvas_mask[None].float(), size=hw).to(torch.bool)[0]


def inverse_sigmoid(x, eps=1e-5):
    """DETR/deformable_detr/util/misc.py:513-517."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def _ln(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


# --------------------------------------------------------------------------------------------------------------
RESNET50_BLOCKS = (3, 4, 6, 3)


def trunk_shapes(prefix="0.body."):
    """Key -> shape table of the reference trunk's state dict: torchvision ResNet-50 behind IntermediateLayerGetter (no
    avgpool / fc), every norm a FrozenBatchNorm2d with four buffers (DETR/deformable_detr/models/backbone.py:25-109)."""
    shapes = {}

    def bn(k, c):
        for f in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + k + "." + f] = (c,)

    shapes[prefix + "conv1.weight"] = (64, 3, 7, 7)
    bn("bn1", 64)
    cin = 64
    for li, nblk in enumerate(RESNET50_BLOCKS):
        mid = 64 * 2 ** li
        for b in range(nblk):
            k = "layer%d.%d." % (li + 1, b)
            shapes[prefix + k + "conv1.weight"] = (mid, cin, 1, 1)
            bn(k + "bn1", mid)
            shapes[prefix + k + "conv2.weight"] = (mid, mid, 3, 3)
            bn(k + "bn2", mid)
            shapes[prefix + k + "conv3.weight"] = (mid * 4, mid, 1, 1)
            bn(k + "bn3", mid * 4)
            if b == 0:
                shapes[prefix + k + "downsample.0.weight"] = (mid * 4, cin, 1, 1)
                bn(k + "downsample.1", mid * 4)
            cin = mid * 4
    return shapes


def trunk_trainable(prefix="0.body."):
    """BackboneBase.__init__ (:72-75): only parameters of layer2-4 train; FrozenBN has buffers only."""
    return sorted(k for k in trunk_shapes(prefix) if k.endswith("weight") and "bn" not in k and "downsample.1" not in k
                  and any(l in k for l in ("layer2", "layer3", "layer4")))


def trunk(sd, x, canvas_mask, prefix="0.body.", eps=1e-5):
    """torchvision-style ResNet-50 (stride on the 3x3, unlike Detectron2's STRIDE_IN_1X1), FrozenBN as x * scale + bias
    with scale = w * rsqrt(var + eps) (:54-64); returns layer2-4 outputs (strides 8 / 16 / 32), their nearest-resized
    padding masks (:91) and sine position embeddings (Joiner.forward, :112-129)."""
    def fbn(k, t):
        scale = sd[prefix + k + ".weight"] * (sd[prefix + k + ".running_var"] + eps).rsqrt()
        bias = sd[prefix + k + ".bias"] - sd[prefix + k + ".running_mean"] * scale
        return t * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)

    t = F.relu(fbn("bn1", F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)))
    t = F.max_pool2d(t, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nblk in enumerate(RESNET50_BLOCKS):
        for b in range(nblk):
            k = "layer%d.%d." % (li + 1, b)
            stride = 2 if (b == 0 and li > 0) else 1
            y = F.relu(fbn(k + "bn1", F.conv2d(t, sd[prefix + k + "conv1.weight"])))
            y = F.relu(fbn(k + "bn2", F.conv2d(y, sd[prefix + k + "conv2.weight"], stride=stride, padding=1)))
            y = fbn(k + "bn3", F.conv2d(y, sd[prefix + k + "conv3.weight"]))
            if b == 0:
                t = fbn(k + "downsample.1", F.conv2d(t, sd[prefix + k + "downsample.0.weight"], stride=stride))
            t = F.relu(y + t)
        if li > 0:
            outs.append(t)
    masks = [level_mask(canvas_mask, o.shape[-2:]) for o in outs]
    pos = [sine_position_embedding(m, 128).to(o.dtype) for m, o in zip(masks, outs)]
    return outs, masks, pos


# --------------------------------------------------------------------------------------------------------------
def ms_deform_attn(sd, key, query, reference_points, inp, shapes, starts, padding_mask, heads, points):
    """DETR/deformable_detr/models/ops/modules/ms_deform_attn.py:78-115."""
    n, lq, c = query.shape
    levels = len(shapes)
    value = _lin(sd, key + ".value_proj", inp)
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(n, -1, heads, c // heads)
    off = _lin(sd, key + ".sampling_offsets", query).view(n, lq, heads, levels, points, 2)
    attn = _lin(sd, key + ".attention_weights", query).view(n, lq, heads, levels * points)
    attn = F.softmax(attn, -1).view(n, lq, heads, levels, points)
    if reference_points.shape[-1] == 2:
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)                  # (W_l, H_l)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / points * reference_points[:, :, None, :, None, 2:] * 0.5
    out = msda_forward(value, shapes, starts, loc, attn)
    return _lin(sd, key + ".output_proj", out)


def encoder_reference_points(shapes, valid_ratios):
    """DETR/deformable_detr/models/deformable_transformer.py:238-250: pixel centres, normalised by the VALID extent of
    each level, then re-scaled by every level's valid ratio -> (N, sum HW, L, 2)."""
    refs = []
    for lvl, (h, w) in enumerate(shapes):
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=torch.float32),
                                torch.linspace(0.5, w - 0.5, w, dtype=torch.float32), indexing="ij")
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        refs.append(torch.stack((rx, ry), -1))
    ref = torch.cat(refs, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def valid_ratio(mask):
    """(w, h) fraction of a level that is not padding (deformable_transformer.py:117-124)."""
    _, h, w = mask.shape
    vh = (~mask[:, :, 0]).sum(1).float() / h
    vw = (~mask[:, 0, :]).sum(1).float() / w
    return torch.stack([vw, vh], -1)


def transformer(sd, cfg, srcs, masks, pos_embeds):
    """DETR/deformable_detr/models/deformable_transformer.py:126-187 (two_stage=False) -> (hs (layers, N, Q, C),
    init_reference, inter_references)."""
    heads, points = cfg["nhead"], cfg["points"]
    shapes = [tuple(s.shape[-2:]) for s in srcs]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    mask = torch.cat([m.flatten(1) for m in masks], 1)
    pos = torch.cat([p.flatten(2).transpose(1, 2) + sd["transformer.level_embed"][lvl].view(1, 1, -1)
                     for lvl, p in enumerate(pos_embeds)], 1)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    ratios = torch.stack([valid_ratio(m) for m in masks], 1)                                 # (N, L, 2)

    # encoder (:189-256): deformable self-attention + FFN, post-norm
    ref = encoder_reference_points(shapes, ratios)
    x = src
    for i in range(cfg["enc_layers"]):
        k = "transformer.encoder.layers.%d" % i
        x = _ln(sd, k + ".norm1", x + ms_deform_attn(sd, k + ".self_attn", x + pos, ref, x, shapes, starts, mask, heads, points))
        x = _ln(sd, k + ".norm2", x + _lin(sd, k + ".linear2", F.relu(_lin(sd, k + ".linear1", x))))
    memory = x

    # decoder (:261-358)
    n, c = memory.shape[0], memory.shape[2]
    query_pos, tgt = torch.split(sd["query_embed.weight"], c, dim=1)
    query_pos, tgt = query_pos[None].expand(n, -1, -1), tgt[None].expand(n, -1, -1)
    reference = _lin(sd, "transformer.reference_points", query_pos).sigmoid()
    init_reference = reference
    hs, inter = [], []
    for i in range(cfg["dec_layers"]):
        k = "transformer.decoder.layers.%d" % i
        ref_in = reference[:, :, None] * ratios[:, None]
        # self-attention between the queries: nn.MultiheadAttention (packed in_proj), q = k = tgt + query_pos, v = tgt
        q = tgt + query_pos
        wq, wk, wv = sd[k + ".self_attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = sd[k + ".self_attn.in_proj_bias"].chunk(3, 0)
        nq = q.shape[1]
        split = lambda t: t.view(n, nq, heads, c // heads).transpose(1, 2)                    # noqa: E731
        a = F.scaled_dot_product_attention(split(F.linear(q, wq, bq)), split(F.linear(q, wk, bk)), split(F.linear(tgt, wv, bv)))
        a = _lin(sd, k + ".self_attn.out_proj", a.transpose(1, 2).reshape(n, nq, c))
        tgt = _ln(sd, k + ".norm2", tgt + a)
        a = ms_deform_attn(sd, k + ".cross_attn", tgt + query_pos, ref_in, memory, shapes, starts, mask, heads, points)
        tgt = _ln(sd, k + ".norm1", tgt + a)
        tgt = _ln(sd, k + ".norm3", tgt + _lin(sd, k + ".linear2", F.relu(_lin(sd, k + ".linear1", tgt))))
        hs.append(tgt)
        inter.append(reference)                       # no box refinement: the reference points never move
    return torch.stack(hs), init_reference, torch.stack(inter)


def forward(sd, cfg, feats, canvas_mask):
    """DETR/deformable_detr/models/deformable_detr.py:119-192 from the trunk's feature maps on."""
    d = cfg["d_model"]
    srcs, masks, pos = [], [], []
    for lvl, f in enumerate(feats):
        k = "input_proj.%d" % lvl
        s = F.group_norm(F.conv2d(f, sd[k + ".0.weight"], sd[k + ".0.bias"]), 32, sd[k + ".1.weight"], sd[k + ".1.bias"])
        m = level_mask(canvas_mask, s.shape[-2:])
        srcs.append(s)
        masks.append(m)
        pos.append(sine_position_embedding(m, d // 2).to(s.dtype))
    for lvl in range(len(feats), cfg["levels"]):      # extra levels: 3x3 stride-2 conv on the RAW last feature map, then on itself
        k = "input_proj.%d" % lvl
        x = feats[-1] if lvl == len(feats) else srcs[-1]
        s = F.group_norm(F.conv2d(x, sd[k + ".0.weight"], sd[k + ".0.bias"], stride=2, padding=1), 32, sd[k + ".1.weight"],
                         sd[k + ".1.bias"])
        m = level_mask(canvas_mask, s.shape[-2:])
        srcs.append(s)
        masks.append(m)
        pos.append(sine_position_embedding(m, d // 2).to(s.dtype))
    hs, init_ref, inter_ref = transformer(sd, cfg, srcs, masks, pos)
    logits, boxes = [], []
    for lvl in range(hs.shape[0]):
        ref = inverse_sigmoid(init_ref if lvl == 0 else inter_ref[lvl - 1])
        h = hs[lvl]
        logits.append(_lin(sd, "class_embed.0", h))
        t = _lin(sd, "bbox_embed.0.layers.2", F.relu(_lin(sd, "bbox_embed.0.layers.1", F.relu(_lin(sd, "bbox_embed.0.layers.0", h)))))
        t = torch.cat([t[..., :2] + ref, t[..., 2:]], -1)
        boxes.append(t.sigmoid())
    return {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
            "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}


# --------------------------------------------------------------------------------------------------------------
def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def generalized_iou(a, b):
    """pairwise GIoU of xyxy boxes (DETR/deformable_detr/util/box_ops.py:32-69)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b - inter
    iou = inter / union
    hull = (torch.max(a[:, None, 2:], b[:, 2:]) - torch.min(a[:, None, :2], b[:, :2])).clamp(min=0)
    hull_area = hull[..., 0] * hull[..., 1]
    return iou - (hull_area - union) / hull_area


@torch.no_grad()
def hungarian(pred_logits, pred_boxes, targets, cost=(2.0, 5.0, 2.0)):
    """DETR/deformable_detr/models/matcher.py:65-96: focal-style class cost (alpha .25, gamma 2, eps 1e-8), L1 and -GIoU
    on the whole batch at once, then one scipy assignment per image on its own block of target columns."""
    n, q = pred_logits.shape[:2]
    prob = pred_logits.flatten(0, 1).sigmoid()
    box = pred_boxes.flatten(0, 1)
    ids = torch.cat([t["labels"] for t in targets])
    tb = torch.cat([t["boxes"] for t in targets])
    neg = 0.75 * prob ** 2.0 * (-(1 - prob + 1e-8).log())
    posc = 0.25 * (1 - prob) ** 2.0 * (-(prob + 1e-8).log())
    c = cost[1] * torch.cdist(box, tb, p=1) + cost[0] * (posc[:, ids] - neg[:, ids]) \
        - cost[2] * generalized_iou(cxcywh_to_xyxy(box), cxcywh_to_xyxy(tb))
    c = c.view(n, q, -1)
    out, col = [], 0
    for i, t in enumerate(targets):
        k = len(t["boxes"])
        r, j = linear_sum_assignment(c[i, :, col:col + k].numpy())
        out.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        col += k
    return out


def focal_loss(logits, onehot, num_boxes, alpha, gamma=2.0):
    """DETR/deformable_detr/models/segmentation.py:196-221: mean over QUERIES (dim 1), sum over the rest, / num_boxes."""
    p = logits.sigmoid()
    ce = F.binary_cross_entropy_with_logits(logits, onehot, reduction="none")
    p_t = p * onehot + (1 - p) * (1 - onehot)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * onehot + (1 - alpha) * (1 - onehot)) * loss
    return loss.mean(1).sum() / num_boxes


def layer_losses(out, targets, indices, num_boxes, num_classes, alpha, log):
    """loss_labels / loss_boxes / loss_cardinality of one decoder layer (deformable_detr.py:218-281)."""
    logits, boxes = out["pred_logits"], out["pred_boxes"]
    bi = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)])
    si = torch.cat([s for s, _ in indices])
    cls_o = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)])
    onehot = torch.zeros_like(logits)
    onehot[bi, si, cls_o] = 1.0
    res = {"loss_ce": focal_loss(logits, onehot, num_boxes, alpha) * logits.shape[1]}
    if log:                                           # top-1 error of the MATCHED queries, in per cent; no gradient
        with torch.no_grad():
            # util/misc.py `accuracy`: correct count * (100 / n); an empty target set gives 0 -> class_error 100
            acc = (logits[bi, si].argmax(-1) == cls_o).float().sum() * (100.0 / cls_o.numel()) if cls_o.numel() else torch.zeros(())
        res["class_error"] = 100.0 - acc
    src = boxes[bi, si]
    tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, indices)], 0)
    res["loss_bbox"] = F.l1_loss(src, tgt, reduction="none").sum() / num_boxes
    res["loss_giou"] = (1 - torch.diag(generalized_iou(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt)))).sum() / num_boxes
    with torch.no_grad():                             # logged only, and multiplied by 0 in the reference (:257)
        lengths = torch.as_tensor([len(t["labels"]) for t in targets], dtype=torch.float32)
        card = (logits.argmax(-1) != logits.shape[-1] - 1).sum(1).float()
        res["cardinality_error"] = F.l1_loss(card, lengths) * 0
    return res


def criterion(out, targets, cfg):
    """SetCriterion.forward (deformable_detr.py:334-370), single process: final layer + every auxiliary layer, each with
    its OWN assignment; num_boxes = total GT in the batch, at least 1.  Returns (unweighted losses, assignments)."""
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)), 1.0)
    ind = [hungarian(out["pred_logits"], out["pred_boxes"], targets, cfg["cost"])]
    losses = layer_losses(out, targets, ind[0], num_boxes, cfg["classes"], cfg["focal_alpha"], True)
    for i, aux in enumerate(out["aux_outputs"]):
        ind.append(hungarian(aux["pred_logits"], aux["pred_boxes"], targets, cfg["cost"]))
        for k, v in layer_losses(aux, targets, ind[-1], num_boxes, cfg["classes"], cfg["focal_alpha"], False).items():
            losses["%s_%d" % (k, i)] = v
    return losses, ind


def weight_dict(cfg):
    """meta_arch.py:109-120: the three weights replicated for every auxiliary layer and `_enc`."""
    w = dict(cfg["weights"])
    for i in range(cfg["dec_layers"] - 1):
        w.update({"%s_%d" % (k, i): v for k, v in cfg["weights"].items()})
    w.update({k + "_enc": v for k, v in cfg["weights"].items()})
    return w


def inference_topk(pred_logits, pred_boxes, image_sizes, topk):
    """meta_arch.py:197-238: top-k over the flattened (query, class) sigmoid scores; no NMS."""
    prob = pred_logits.sigmoid()
    scores, idx = torch.topk(prob.view(prob.shape[0], -1), topk, dim=1)
    q = torch.div(idx, prob.shape[2], rounding_mode="trunc")
    labels = idx % prob.shape[2]
    boxes = cxcywh_to_xyxy(torch.gather(pred_boxes, 1, q.unsqueeze(-1).repeat(1, 1, 4)))
    scale = torch.tensor([[w, h, w, h] for h, w in image_sizes], dtype=boxes.dtype)
    return scores, idx, labels, boxes * scale[:, None]
