"""Parameter table of the Faster R-CNN R50-FPN detector the ALDI++ step trains, with Detectron2 state_dict
key names (so released checkpoints / `DetectionCheckpointerWithEMA` files map 1:1 — aldi/checkpoint.py:8-31,
aldi/ema.py:19-50 iterate over exactly these keys), plus a deterministic synthetic initialiser used by the
benchmark and the parity tests (there is no network for the COCO-pretrained `model_final_f10217.pkl`).
"""
from collections import OrderedDict

import torch


class ConvSpec:
    """One conv / linear layer: Detectron2 key prefix and geometry."""

    def __init__(self, key, cin, cout, k=1, stride=1, pad=0, bias=False, norm=False, trainable=True, linear=False, ln=False):
        self.key, self.cin, self.cout, self.k, self.stride, self.pad = key, cin, cout, k, stride, pad
        self.bias, self.norm, self.trainable, self.linear = bias, norm, trainable, linear
        self.ln = ln      # a trainable channel LayerNorm follows (detectron2 get_norm("LN")): `<key>.norm.{weight,bias}`


RES_DEPTHS = (3, 4, 6, 3)


def resnet_specs(freeze_at=2):
    specs = OrderedDict()
    bu = "backbone.bottom_up."
    specs["stem"] = ConvSpec(bu + "stem.conv1", 3, 64, 7, 2, 3, norm=True, trainable=freeze_at < 1)
    cin, cout, mid = 64, 256, 64
    for si, nblk in enumerate(RES_DEPTHS):
        stage = si + 2
        tr = freeze_at < stage
        for b in range(nblk):
            stride = 2 if (b == 0 and si > 0) else 1
            p = "%sres%d.%d." % (bu, stage, b)
            name = "res%d.%d." % (stage, b)
            if b == 0:
                specs[name + "shortcut"] = ConvSpec(p + "shortcut", cin, cout, 1, stride, 0, norm=True, trainable=tr)
            specs[name + "conv1"] = ConvSpec(p + "conv1", cin, mid, 1, stride, 0, norm=True, trainable=tr)
            specs[name + "conv2"] = ConvSpec(p + "conv2", mid, mid, 3, 1, 1, norm=True, trainable=tr)
            specs[name + "conv3"] = ConvSpec(p + "conv3", mid, cout, 1, 1, 0, norm=True, trainable=tr)
            cin = cout
        cout *= 2
        mid *= 2
    return specs


def align_specs(align):
    """Discriminators of the AlignMixin (aldi/align.py:40-41,103-136) with their nn.Sequential state_dict keys.
    align: {"img": (input_dim, hidden_dims) | None, "ins": (input_dim, hidden_dims) | None}."""
    specs = OrderedDict()
    if not align:
        return specs
    if align.get("img"):
        cin, hidden = align["img"]
        for i, dim in enumerate(hidden):       # Conv2d(prev, dim, 3) [valid], ReLU
            specs["img_align.conv%d" % i] = ConvSpec("img_align.model.%d" % (2 * i), cin, dim, 3, 1, 0, bias=True)
            cin = dim
        # ..., AdaptiveAvgPool2d(1), Flatten, Linear(prev, 1)
        specs["img_align.out"] = ConvSpec("img_align.model.%d" % (2 * len(hidden) + 2), cin, 1, bias=True, linear=True)
    if align.get("ins"):
        cin, hidden = align["ins"]
        for i, dim in enumerate(hidden):       # Flatten, then Linear(prev, dim), ReLU
            specs["ins_align.fc%d" % i] = ConvSpec("ins_align.model.%d" % (1 + 2 * i), cin, dim, bias=True, linear=True)
            cin = dim
        specs["ins_align.out"] = ConvSpec("ins_align.model.%d" % (1 + 2 * len(hidden)), cin, 1, bias=True, linear=True)
    return specs


VITDET_HEADS = dict(pyramid=True, rpn_convs=2, box_convs=4, box_fcs=1)   # configs/Base-RCNN-VitDetB.yaml:7-14


def rcnn_specs(num_classes=8, freeze_at=2, align=None, bottom_up_channels=None, pyramid=False, rpn_convs=1, box_convs=0,
               box_fcs=2):
    """bottom_up_channels: None -> ResNet-50 bottom-up in this table; a 4-tuple -> another bottom-up (ConvNeXt:
    aldi_b200/convnext.py keeps its own parameter buffer) whose stage widths feed the FPN laterals.
    pyramid: the backbone returns p2..p5 itself (ViTDet's SimpleFeaturePyramid, aldi_b200/vit.py): no ResNet, no FPN here.
    rpn_convs / box_convs / box_fcs: MODEL.RPN.CONV_DIMS ([-1] * k: k 3x3 convs, named conv | conv.conv{i} as
    StandardRPNHead does) and MODEL.ROI_BOX_HEAD.NUM_CONV (3x3 256 + "LN") / NUM_FC of FastRCNNConvFCHead."""
    specs = resnet_specs(freeze_at) if (bottom_up_channels is None and not pyramid) else OrderedDict()
    if not pyramid:
        for lvl, c in zip((2, 3, 4, 5), bottom_up_channels or (256, 512, 1024, 2048)):
            specs["fpn_lateral%d" % lvl] = ConvSpec("backbone.fpn_lateral%d" % lvl, c, 256, 1, 1, 0, bias=True)
            specs["fpn_output%d" % lvl] = ConvSpec("backbone.fpn_output%d" % lvl, 256, 256, 3, 1, 1, bias=True)
    rp = "proposal_generator.rpn_head."
    if rpn_convs == 1:
        specs["rpn_conv"] = ConvSpec(rp + "conv", 256, 256, 3, 1, 1, bias=True)
    else:
        for k in range(rpn_convs):
            specs["rpn_conv%d" % k] = ConvSpec(rp + "conv.conv%d" % k, 256, 256, 3, 1, 1, bias=True)
    specs["rpn_obj"] = ConvSpec(rp + "objectness_logits", 256, 3, 1, 1, 0, bias=True)
    specs["rpn_delta"] = ConvSpec(rp + "anchor_deltas", 256, 12, 1, 1, 0, bias=True)
    for k in range(box_convs):
        specs["box_conv%d" % (k + 1)] = ConvSpec("roi_heads.box_head.conv%d" % (k + 1), 256, 256, 3, 1, 1, ln=True)
    assert box_fcs in (1, 2)
    specs["fc1"] = ConvSpec("roi_heads.box_head.fc1", 256 * 7 * 7, 1024, bias=True)
    if box_fcs == 2:
        specs["fc2"] = ConvSpec("roi_heads.box_head.fc2", 1024, 1024, bias=True)
    specs["cls_score"] = ConvSpec("roi_heads.box_predictor.cls_score", 1024, num_classes + 1, bias=True)
    specs["bbox_pred"] = ConvSpec("roi_heads.box_predictor.bbox_pred", 1024, num_classes * 4, bias=True)
    specs.update(align_specs(align))
    return specs


LINEAR_LAYERS = ("fc1", "fc2", "cls_score", "bbox_pred")
NORM_FIELDS = ("weight", "bias", "running_mean", "running_var")


def d2_shape(name, s):
    """Shape of `<key>.weight` in the Detectron2 state_dict."""
    if name in LINEAR_LAYERS or s.linear:
        return (s.cout, s.cin)
    return (s.cout, s.cin, s.k, s.k)


def state_dict_entries(num_classes=8, freeze_at=2, align=None, bottom_up_channels=None, **head):
    """[(d2_key, shape, layer_name, field)] in a fixed order; field in {weight,bias,norm.<f>}."""
    out = []
    for name, s in rcnn_specs(num_classes, freeze_at, align, bottom_up_channels, **head).items():
        out.append((s.key + ".weight", d2_shape(name, s), name, "weight"))
        if s.bias:
            out.append((s.key + ".bias", (s.cout,), name, "bias"))
        if s.ln:
            out.append((s.key + ".norm.weight", (s.cout,), name, "norm.weight"))
            out.append((s.key + ".norm.bias", (s.cout,), name, "norm.bias"))
        if s.norm:
            for f in NORM_FIELDS:
                out.append(("%s.norm.%s" % (s.key, f), (s.cout,), name, "norm." + f))
    return out


def synthetic_state_dict(seed=0, num_classes=8, freeze_at=2, align=None, bottom_up_channels=None, **head):
    """Deterministic CPU-generated weights with O(1) activations (stands in for a burn-in checkpoint).

    Scales are chosen so that a bf16 trunk stays in range, RPN logits are well separated and the
    box classifier is confident enough (>0.8) on some RoIs for the pseudo-label path to be non-empty.
    """
    sd = OrderedDict()
    for idx, (key, shape, name, field) in enumerate(state_dict_entries(num_classes, freeze_at, align, bottom_up_channels, **head)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if field == "weight":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            fan_out = shape[0] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            if name == "stem" or name.startswith("res"):
                t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
            elif name.startswith("fpn") or name.startswith("box_conv") or name in ("fc1", "fc2") or "_align." in name:
                bound = (3.0 / fan_in) ** 0.5
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            elif name.startswith("rpn"):
                t = torch.randn(shape, generator=g) * 0.03
            elif name == "cls_score":
                t = torch.randn(shape, generator=g) * 0.2
            else:  # bbox_pred
                t = torch.randn(shape, generator=g) * 0.02
        elif field == "bias":
            t = torch.randn(shape, generator=g) * 0.01
        elif field == "norm.weight" and name.startswith("box_conv"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif field == "norm.weight":
            base = 0.05 if name == "stem" else 0.4 if name.endswith("conv3") else 0.8 if name.endswith("shortcut") else 1.0
            t = base * (1.0 + 0.1 * torch.randn(shape, generator=g))
        elif field == "norm.bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif field == "norm.running_mean":
            t = 0.05 * torch.randn(shape, generator=g)
        else:  # running_var
            t = 1.0 + 0.2 * torch.rand(shape, generator=g)
        sd[key] = t.float()
    return sd
