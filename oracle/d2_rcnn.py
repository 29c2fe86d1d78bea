"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the Detectron2 Faster R-CNN
R50-FPN train / inference path that justinkay/aldi delegates to.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

PARITY STATUS: the algorithm lives in a third-party dependency that is ABSENT from /root/reference:
`detectron2 @ git+https://github.com/justinkay/detectron2_v07ish.git` (pyproject.toml:18, no commit pin; a
fork of facebookresearch/detectron2 ~v0.6).  This file restates detectron2 v0.6's published algorithm
(detectron2/modeling/**, detectron2/structures/**, detectron2/layers/**) on the primitives Detectron2
itself dispatches to and that ARE installed here (torch.nn.functional, torchvision.ops.{nms,batched_nms,
roi_align}).  The reference holds no golden vectors for this path: the Detectron2
layer of the oracle is "parity unpinned"; the ALDI layer above it (oracle/aldi_ref.py) is pinned against the
real reference code run in this container (tests/golden/make_golden.py).

Module / parameter names follow Detectron2 so that state_dict keys match released checkpoints
(`backbone.bottom_up.res2.0.conv1.weight`, `backbone.fpn_lateral2.weight`,
`proposal_generator.rpn_head.conv.weight`, `roi_heads.box_head.fc1.weight`, ...), which is what
aldi/ema.py:19-50 and aldi/checkpoint.py:8-31 iterate over.
"""
import math
from typing import Tuple

import torch
import torch.nn.functional as F
import torchvision
from torch import nn

# ---------------------------------------------------------------------------------------------------
# structures (detectron2/structures/boxes.py, instances.py, image_list.py)
# ---------------------------------------------------------------------------------------------------


class Boxes:
    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def area(self):
        box = self.tensor
        return (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])

    def clip(self, box_size):
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold: float = 0.0):
        box = self.tensor
        widths = box[:, 2] - box[:, 0]
        heights = box[:, 3] - box[:, 1]
        return (widths > threshold) & (heights > threshold)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2
        return Boxes(b)

    def __len__(self):
        return self.tensor.shape[0]

    @classmethod
    def cat(cls, boxes_list):
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    @property
    def device(self):
        return self.tensor.device


def pairwise_intersection(boxes1: Boxes, boxes2: Boxes):
    b1, b2 = boxes1.tensor, boxes2.tensor
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    return wh.prod(dim=2)


def pairwise_iou(boxes1: Boxes, boxes2: Boxes):
    area1, area2 = boxes1.area(), boxes2.area()
    inter = pairwise_intersection(boxes1, boxes2)
    return torch.where(inter > 0, inter / (area1[:, None] + area2 - inter), torch.zeros(1, dtype=inter.dtype))


class Instances:
    def __init__(self, image_size: Tuple[int, int], **kwargs):
        self._image_size = image_size
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(data_len, len(self))
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def to(self, *args, **kwargs):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item):
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            item = slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    @staticmethod
    def cat(instance_lists):
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        image_size = instance_lists[0].image_size
        ret = Instances(image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
            ret.set(k, values)
        return ret


class ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
        max_h = max(s[0] for s in image_sizes)
        max_w = max(s[1] for s in image_sizes)
        if size_divisibility > 1:
            st = size_divisibility
            max_h = (max_h + (st - 1)) // st * st
            max_w = (max_w + (st - 1)) // st * st
        batched = tensors[0].new_full((len(tensors), tensors[0].shape[0], max_h, max_w), pad_value)
        for img, pad_img in zip(tensors, batched):
            pad_img[..., : img.shape[-2], : img.shape[-1]].copy_(img)
        return ImageList(batched.contiguous(), image_sizes)


class EventStorage:
    """Stand-in for detectron2.utils.events.EventStorage: records the scalars D2 writes during a training forward."""
    _current = []

    def __init__(self, start_iter=0):
        self.iter = start_iter
        self.scalars = {}

    def put_scalar(self, name, value, smoothing_hint=True):
        self.scalars[name] = float(value)

    def __enter__(self):
        EventStorage._current.append(self)
        return self

    def __exit__(self, *a):
        EventStorage._current.pop()


def get_event_storage():
    if not EventStorage._current:
        # detectron2 asserts here; the oracle is lenient so single-op tests need no context
        EventStorage._current.append(EventStorage())
    return EventStorage._current[-1]


def nonzero_tuple(x):
    return x.nonzero(as_tuple=True)


def cat(tensors, dim=0):
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    """detectron2.layers.wrappers.cross_entropy: 0 (not nan) on empty input."""
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


def smooth_l1_loss(input, target, beta: float, reduction: str = "none"):
    """fvcore.nn.smooth_l1_loss."""
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2.layers.nms.batched_nms == torchvision batched_nms on float boxes."""
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


# ---------------------------------------------------------------------------------------------------
# box regression / matcher / sampling
# ---------------------------------------------------------------------------------------------------
_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1 = pred_ctr_x - 0.5 * pred_w
        y1 = pred_ctr_y - 0.5 * pred_h
        x2 = pred_ctr_x + 0.5 * pred_w
        y2 = pred_ctr_y + 0.5 * pred_h
        pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
        return pred_boxes.reshape(deltas.shape)


def _dense_box_regression_loss(anchors, box2box_transform, pred_anchor_deltas, gt_boxes, fg_mask,
                               box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0):
    if isinstance(anchors[0], Boxes):
        anchors = Boxes.cat(anchors).tensor
    else:
        anchors = cat(anchors)
    assert box_reg_loss_type == "smooth_l1"
    gt_anchor_deltas = [box2box_transform.get_deltas(anchors, k) for k in gt_boxes]
    gt_anchor_deltas = torch.stack(gt_anchor_deltas)
    return smooth_l1_loss(cat(pred_anchor_deltas, dim=1)[fg_mask], gt_anchor_deltas[fg_mask], beta=smooth_l1_beta,
                          reduction="sum")


class Matcher:
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix):
        assert match_quality_matrix.dim() == 2
        if match_quality_matrix.numel() == 0:
            default_matches = match_quality_matrix.new_full((match_quality_matrix.size(1),), 0, dtype=torch.int64)
            default_match_labels = match_quality_matrix.new_full((match_quality_matrix.size(1),), self.labels[0],
                                                                 dtype=torch.int8)
            return default_matches, default_match_labels
        assert torch.all(match_quality_matrix >= 0)
        matched_vals, matches = match_quality_matrix.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        if self.allow_low_quality_matches:
            self.set_low_quality_matches_(match_labels, match_quality_matrix)
        return matches, match_labels

    def set_low_quality_matches_(self, match_labels, match_quality_matrix):
        highest_quality_foreach_gt, _ = match_quality_matrix.max(dim=1)
        _, pred_inds_with_highest_quality = nonzero_tuple(match_quality_matrix == highest_quality_foreach_gt[:, None])
        match_labels[pred_inds_with_highest_quality] = 1


# Sampler hook.  Reference behaviour: torch.randperm on the model device under the global generator
# (detectron2/modeling/sampling.py); CPU and CUDA randperm streams differ, so "identical inputs" parity
# needs an injectable, platform-independent choice (SURVEY.md §7 hard parts, T3).  `None` = reference
# behaviour; tests install a callable (num_candidates, num_take, tag) -> LongTensor of positions.
_CHOOSER = None
# where in the step the current subsample_labels call sits (only read by an installed chooser)
SAMPLE_CTX = {"pass": 0, "site": None, "site_override": None, "image": 0}


def set_sample_chooser(fn):
    """fn(candidate_indices: LongTensor, take: int, tag: 'pos'|'neg', ctx: dict) -> LongTensor of positions."""
    global _CHOOSER
    _CHOOSER = fn


def _choose(candidates, take, tag):
    if _CHOOSER is not None:
        return _CHOOSER(candidates, take, tag, dict(SAMPLE_CTX)).to(candidates.device)
    return torch.randperm(candidates.numel(), device=candidates.device)[:take]


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = _choose(positive, num_pos, "pos")
    perm2 = _choose(negative, num_neg, "neg")
    return positive[perm1], negative[perm2]


# ---------------------------------------------------------------------------------------------------
# backbone: ResNet-50 (detectron2/modeling/backbone/resnet.py) + FPN (fpn.py)
# ---------------------------------------------------------------------------------------------------
class FrozenBatchNorm2d(nn.Module):
    _version = 3

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        if x.requires_grad:
            scale = self.weight * (self.running_var + self.eps).rsqrt()
            bias = self.bias - self.running_mean * scale
            scale = scale.reshape(1, -1, 1, 1)
            bias = bias.reshape(1, -1, 1, 1)
            out_dtype = x.dtype
            return x * scale.to(out_dtype) + bias.to(out_dtype)
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, training=False,
                            eps=self.eps)


class Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv -> norm -> activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def c2_msra_fill(module):
    nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


class BasicStem(nn.Module):
    def __init__(self, in_channels=3, out_channels=64):
        super().__init__()
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=7, stride=2, padding=3, bias=False,
                            norm=FrozenBatchNorm2d(out_channels))
        c2_msra_fill(self.conv1)

    def forward(self, x):
        x = self.conv1(x)
        x = F.relu_(x)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        return x


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, stride_in_1x1=True):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=FrozenBatchNorm2d(out_channels))
        else:
            self.shortcut = None
        stride_1x1, stride_3x3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=stride_1x1, bias=False,
                            norm=FrozenBatchNorm2d(bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=stride_3x3, padding=1,
                            bias=False, norm=FrozenBatchNorm2d(bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False,
                            norm=FrozenBatchNorm2d(out_channels))
        for layer in [self.conv1, self.conv2, self.conv3, self.shortcut]:
            if layer is not None:
                c2_msra_fill(layer)

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        shortcut = self.shortcut(x) if self.shortcut is not None else x
        out += shortcut
        out = F.relu_(out)
        return out

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        return self


class ResNet(nn.Module):
    def __init__(self, freeze_at=2, depths=(3, 4, 6, 3)):
        super().__init__()
        self.stem = BasicStem()
        self.stage_names = []
        in_ch, out_ch, bott = 64, 256, 64
        for idx, nblk in enumerate(depths):
            first_stride = 1 if idx == 0 else 2
            blocks = []
            for i in range(nblk):
                blocks.append(BottleneckBlock(in_ch, out_ch, bottleneck_channels=bott,
                                              stride=first_stride if i == 0 else 1, stride_in_1x1=True))
                in_ch = out_ch
            name = "res%d" % (idx + 2)
            self.add_module(name, nn.Sequential(*blocks))
            self.stage_names.append(name)
            out_ch *= 2
            bott *= 2
        self._out_feature_channels = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
        self._out_feature_strides = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}
        self.freeze(freeze_at)

    def freeze(self, freeze_at=0):
        if freeze_at >= 1:
            for p in self.stem.parameters():
                p.requires_grad = False
        for idx, name in enumerate(self.stage_names, start=2):
            if freeze_at >= idx:
                for block in getattr(self, name).children():
                    block.freeze()
        return self

    def forward(self, x):
        outputs = {}
        x = self.stem(x)
        for name in self.stage_names:
            x = getattr(self, name)(x)
            outputs[name] = x
        return outputs


class FPN(nn.Module):
    def __init__(self, bottom_up, in_features=("res2", "res3", "res4", "res5"), out_channels=256):
        super().__init__()
        self.bottom_up = bottom_up
        self.in_features = tuple(in_features)
        strides = [bottom_up._out_feature_strides[f] for f in in_features]
        in_channels_per_feature = [bottom_up._out_feature_channels[f] for f in in_features]
        lateral_convs, output_convs = [], []
        for idx, in_channels in enumerate(in_channels_per_feature):
            lateral_conv = Conv2d(in_channels, out_channels, kernel_size=1, bias=True)
            output_conv = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=True)
            c2_xavier_fill(lateral_conv)
            c2_xavier_fill(output_conv)
            stage = int(math.log2(strides[idx]))
            self.add_module("fpn_lateral{}".format(stage), lateral_conv)
            self.add_module("fpn_output{}".format(stage), output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs[::-1]
        self._out_features = ["p2", "p3", "p4", "p5", "p6"]
        self._out_feature_strides = {"p2": 4, "p3": 8, "p4": 16, "p5": 32, "p6": 64}
        self.size_divisibility = 32

    def forward(self, x):
        bottom_up_features = self.bottom_up(x)
        results = []
        prev_features = self.lateral_convs[0](bottom_up_features[self.in_features[-1]])
        results.append(self.output_convs[0](prev_features))
        for idx, (lateral_conv, output_conv) in enumerate(zip(self.lateral_convs, self.output_convs)):
            if idx > 0:
                features = bottom_up_features[self.in_features[-idx - 1]]
                top_down_features = F.interpolate(prev_features, scale_factor=2.0, mode="nearest")
                lateral_features = lateral_conv(features)
                prev_features = lateral_features + top_down_features
                results.insert(0, output_conv(prev_features))
        # LastLevelMaxPool on p5
        results.append(F.max_pool2d(results[3], kernel_size=1, stride=2, padding=0))
        return {f: res for f, res in zip(self._out_features, results)}


# ---------------------------------------------------------------------------------------------------
# RPN (detectron2/modeling/anchor_generator.py, proposal_generator/rpn.py, proposal_utils.py)
# ---------------------------------------------------------------------------------------------------
class BufferList(nn.Module):
    def __init__(self, buffers):
        super().__init__()
        for i, buffer in enumerate(buffers):
            self.register_buffer(str(i), buffer, persistent=False)

    def __len__(self):
        return len(self._buffers)

    def __iter__(self):
        return iter(self._buffers.values())


class DefaultAnchorGenerator(nn.Module):
    box_dim = 4

    def __init__(self, sizes=((32,), (64,), (128,), (256,), (512,)), aspect_ratios=((0.5, 1.0, 2.0),),
                 strides=(4, 8, 16, 32, 64), offset=0.0):
        super().__init__()
        self.strides = strides
        self.num_features = len(strides)
        if len(aspect_ratios) == 1:
            aspect_ratios = list(aspect_ratios) * self.num_features
        self.cell_anchors = BufferList([self.generate_cell_anchors(s, a).float() for s, a in zip(sizes, aspect_ratios)])
        self.offset = offset

    @property
    def num_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @staticmethod
    def generate_cell_anchors(sizes, aspect_ratios):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for aspect_ratio in aspect_ratios:
                w = math.sqrt(area / aspect_ratio)
                h = aspect_ratio * w
                x0, y0, x1, y1 = -w / 2.0, -h / 2.0, w / 2.0, h / 2.0
                anchors.append([x0, y0, x1, y1])
        return torch.tensor(anchors)

    def _grid_anchors(self, grid_sizes):
        anchors = []
        for size, stride, base_anchors in zip(grid_sizes, self.strides, self.cell_anchors):
            grid_height, grid_width = size
            shifts_x = torch.arange(self.offset * stride, grid_width * stride, step=stride, dtype=torch.float32)
            shifts_y = torch.arange(self.offset * stride, grid_height * stride, step=stride, dtype=torch.float32)
            shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
            shift_x = shift_x.reshape(-1)
            shift_y = shift_y.reshape(-1)
            shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
            anchors.append((shifts.view(-1, 1, 4) + base_anchors.view(1, -1, 4)).reshape(-1, 4))
        return anchors

    def forward(self, features):
        grid_sizes = [feature_map.shape[-2:] for feature_map in features]
        return [Boxes(x) for x in self._grid_anchors(grid_sizes)]


class StandardRPNHead(nn.Module):
    def __init__(self, in_channels=256, num_anchors=3, box_dim=4, conv_dims=(-1,)):
        """conv_dims: MODEL.RPN.CONV_DIMS (-1 = the input width).  One entry: the 3x3 conv is `conv`; several (ViTDet's
        [-1, -1], configs/Base-RCNN-VitDetB.yaml:13-14): `conv` is a Sequential of `conv0`, `conv1`, ... each + ReLU."""
        super().__init__()
        cur = in_channels
        if len(conv_dims) == 1:
            out = cur if conv_dims[0] == -1 else conv_dims[0]
            self.conv = Conv2d(cur, out, kernel_size=3, stride=1, padding=1, activation=nn.ReLU())
            convs, cur = [self.conv], out
        else:
            self.conv = nn.Sequential()
            convs = []
            for k, d in enumerate(conv_dims):
                out = cur if d == -1 else d
                c = Conv2d(cur, out, kernel_size=3, stride=1, padding=1, activation=nn.ReLU())
                self.conv.add_module("conv%d" % k, c)
                convs.append(c)
                cur = out
        self.objectness_logits = nn.Conv2d(cur, num_anchors, kernel_size=1, stride=1)
        self.anchor_deltas = nn.Conv2d(cur, num_anchors * box_dim, kernel_size=1, stride=1)
        for layer in convs + [self.objectness_logits, self.anchor_deltas]:
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)

    def forward(self, features):
        pred_objectness_logits, pred_anchor_deltas = [], []
        for x in features:
            t = self.conv(x)
            pred_objectness_logits.append(self.objectness_logits(t))
            pred_anchor_deltas.append(self.anchor_deltas(t))
        return pred_objectness_logits, pred_anchor_deltas


def find_top_rpn_proposals(proposals, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_size, training):
    num_images = len(image_sizes)
    device = proposals[0].device
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images, device=device)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        Hi_Wi_A = logits_i.shape[1]
        num_proposals_i = min(Hi_Wi_A, pre_nms_topk)
        topk_scores_i, topk_idx = logits_i.topk(num_proposals_i, dim=1)
        topk_proposals_i = proposals_i[batch_idx[:, None], topk_idx]
        topk_proposals.append(topk_proposals_i)
        topk_scores.append(topk_scores_i)
        level_ids.append(torch.full((num_proposals_i,), level_id, dtype=torch.int64, device=device))
    topk_scores = cat(topk_scores, dim=1)
    topk_proposals = cat(topk_proposals, dim=1)
    level_ids = cat(level_ids, dim=0)
    results = []
    for n, image_size in enumerate(image_sizes):
        boxes = Boxes(topk_proposals[n])
        scores_per_img = topk_scores[n]
        lvl = level_ids
        valid_mask = torch.isfinite(boxes.tensor).all(dim=1) & torch.isfinite(scores_per_img)
        if not valid_mask.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes = boxes[valid_mask]
            scores_per_img = scores_per_img[valid_mask]
            lvl = lvl[valid_mask]
        boxes.clip(image_size)
        keep = boxes.nonempty(threshold=min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores_per_img, lvl = boxes[keep], scores_per_img[keep], lvl[keep]
        keep = batched_nms(boxes.tensor, scores_per_img, lvl, nms_thresh)
        keep = keep[:post_nms_topk]
        res = Instances(image_size)
        res.proposal_boxes = boxes[keep]
        res.objectness_logits = scores_per_img[keep]
        results.append(res)
    return results


class RPN(nn.Module):
    def __init__(self, in_features=("p2", "p3", "p4", "p5", "p6"), batch_size_per_image=256, positive_fraction=0.5,
                 pre_nms_topk=(2000, 1000), post_nms_topk=(1000, 1000), nms_thresh=0.7, min_box_size=0.0,
                 iou_thresholds=(0.3, 0.7), iou_labels=(0, -1, 1), smooth_l1_beta=0.0, conv_dims=(-1,)):
        super().__init__()
        self.in_features = in_features
        self.rpn_head = StandardRPNHead(conv_dims=conv_dims)
        self.anchor_generator = DefaultAnchorGenerator()
        self.anchor_matcher = Matcher(list(iou_thresholds), list(iou_labels), allow_low_quality_matches=True)
        self.box2box_transform = Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0))
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.pre_nms_topk = {True: pre_nms_topk[0], False: pre_nms_topk[1]}
        self.post_nms_topk = {True: post_nms_topk[0], False: post_nms_topk[1]}
        self.nms_thresh = nms_thresh
        self.min_box_size = float(min_box_size)
        self.smooth_l1_beta = smooth_l1_beta
        self.loss_weight = {"loss_rpn_cls": 1.0, "loss_rpn_loc": 1.0}

    def _subsample_labels(self, label):
        pos_idx, neg_idx = subsample_labels(label, self.batch_size_per_image, self.positive_fraction, 0)
        label.fill_(-1)
        label.scatter_(0, pos_idx, 1)
        label.scatter_(0, neg_idx, 0)
        return label

    @torch.no_grad()
    def label_and_sample_anchors(self, anchors, gt_instances):
        anchors = Boxes.cat(anchors)
        gt_boxes = [x.gt_boxes for x in gt_instances]
        gt_labels, matched_gt_boxes = [], []
        for i, gt_boxes_i in enumerate(gt_boxes):
            SAMPLE_CTX["site"], SAMPLE_CTX["image"] = SAMPLE_CTX["site_override"] or "rpn", i
            match_quality_matrix = pairwise_iou(gt_boxes_i, anchors)
            matched_idxs, gt_labels_i = self.anchor_matcher(match_quality_matrix)
            gt_labels_i = gt_labels_i.to(device=gt_boxes_i.device)
            del match_quality_matrix
            gt_labels_i = self._subsample_labels(gt_labels_i)
            if len(gt_boxes_i) == 0:
                matched_gt_boxes_i = torch.zeros_like(anchors.tensor)
            else:
                matched_gt_boxes_i = gt_boxes_i[matched_idxs].tensor
            gt_labels.append(gt_labels_i)
            matched_gt_boxes.append(matched_gt_boxes_i)
        return gt_labels, matched_gt_boxes

    def losses(self, anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes):
        num_images = len(gt_labels)
        gt_labels = torch.stack(gt_labels)
        pos_mask = gt_labels == 1
        num_pos_anchors = pos_mask.sum().item()
        num_neg_anchors = (gt_labels == 0).sum().item()
        storage = get_event_storage()
        storage.put_scalar("rpn/num_pos_anchors", num_pos_anchors / num_images)
        storage.put_scalar("rpn/num_neg_anchors", num_neg_anchors / num_images)
        localization_loss = _dense_box_regression_loss(anchors, self.box2box_transform, pred_anchor_deltas, gt_boxes,
                                                       pos_mask, smooth_l1_beta=self.smooth_l1_beta)
        valid_mask = gt_labels >= 0
        objectness_loss = F.binary_cross_entropy_with_logits(cat(pred_objectness_logits, dim=1)[valid_mask],
                                                             gt_labels[valid_mask].to(torch.float32), reduction="sum")
        normalizer = self.batch_size_per_image * num_images
        losses = {"loss_rpn_cls": objectness_loss / normalizer, "loss_rpn_loc": localization_loss / normalizer}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def forward(self, images, features, gt_instances=None):
        features = [features[f] for f in self.in_features]
        anchors = self.anchor_generator(features)
        pred_objectness_logits, pred_anchor_deltas = self.rpn_head(features)
        pred_objectness_logits = [score.permute(0, 2, 3, 1).flatten(1) for score in pred_objectness_logits]
        pred_anchor_deltas = [
            x.view(x.shape[0], -1, self.anchor_generator.box_dim, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2)
            for x in pred_anchor_deltas
        ]
        if self.training:
            assert gt_instances is not None, "RPN requires gt_instances in training!"
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes)
        else:
            losses = {}
        proposals = self.predict_proposals(anchors, pred_objectness_logits, pred_anchor_deltas, images.image_sizes)
        return proposals, losses

    def predict_proposals(self, anchors, pred_objectness_logits, pred_anchor_deltas, image_sizes):
        with torch.no_grad():
            pred_proposals = self._decode_proposals(anchors, pred_anchor_deltas)
            return find_top_rpn_proposals(pred_proposals, pred_objectness_logits, image_sizes, self.nms_thresh,
                                          self.pre_nms_topk[self.training], self.post_nms_topk[self.training],
                                          self.min_box_size, self.training)

    def _decode_proposals(self, anchors, pred_anchor_deltas):
        N = pred_anchor_deltas[0].shape[0]
        proposals = []
        for anchors_i, pred_anchor_deltas_i in zip(anchors, pred_anchor_deltas):
            B = anchors_i.tensor.size(1)
            pred_anchor_deltas_i = pred_anchor_deltas_i.reshape(-1, B)
            anchors_i = anchors_i.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, B)
            proposals_i = self.box2box_transform.apply_deltas(pred_anchor_deltas_i, anchors_i)
            proposals.append(proposals_i.view(N, -1, B))
        return proposals


# ---------------------------------------------------------------------------------------------------
# ROI heads (detectron2/modeling/poolers.py, roi_heads/{roi_heads,box_head,fast_rcnn}.py)
# ---------------------------------------------------------------------------------------------------
def assign_boxes_to_levels(box_lists, min_level, max_level, canonical_box_size, canonical_level):
    box_sizes = torch.sqrt(cat([boxes.area() for boxes in box_lists]))
    level_assignments = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    level_assignments = torch.clamp(level_assignments, min=min_level, max=max_level)
    return level_assignments.to(torch.int64) - min_level


def convert_boxes_to_pooler_format(box_lists):
    boxes = torch.cat([x.tensor for x in box_lists], dim=0)
    sizes = torch.tensor([len(x) for x in box_lists])
    indices = torch.repeat_interleave(torch.arange(len(box_lists), dtype=boxes.dtype), sizes)
    return cat([indices[:, None], boxes], dim=1)


class ROIPooler(nn.Module):
    def __init__(self, output_size=7, scales=(1 / 4, 1 / 8, 1 / 16, 1 / 32), sampling_ratio=0, canonical_box_size=224,
                 canonical_level=4):
        super().__init__()
        self.output_size = (output_size, output_size)
        self.scales = scales
        self.sampling_ratio = sampling_ratio
        self.min_level = int(-math.log2(scales[0]))
        self.max_level = int(-math.log2(scales[-1]))
        self.canonical_level = canonical_level
        self.canonical_box_size = canonical_box_size

    def forward(self, x, box_lists):
        num_level_assignments = len(self.scales)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        level_assignments = assign_boxes_to_levels(box_lists, self.min_level, self.max_level, self.canonical_box_size,
                                                   self.canonical_level)
        num_boxes = pooler_fmt_boxes.size(0)
        num_channels = x[0].shape[1]
        output = torch.zeros((num_boxes, num_channels, self.output_size[0], self.output_size[1]), dtype=x[0].dtype,
                             device=x[0].device)
        for level in range(num_level_assignments):
            inds = nonzero_tuple(level_assignments == level)[0]
            pooler_fmt_boxes_level = pooler_fmt_boxes[inds]
            pooled = torchvision.ops.roi_align(x[level], pooler_fmt_boxes_level.to(dtype=x[level].dtype),
                                               self.output_size, self.scales[level], self.sampling_ratio, aligned=True)
            output.index_put_((inds,), pooled)
        return output


class LayerNorm2d(nn.Module):
    """detectron2.layers.batch_norm.LayerNorm (`get_norm("LN", c)`): over the channel axis of NCHW, eps 1e-6."""

    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight, self.bias, self.eps = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c)), eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class FastRCNNConvFCHead(nn.Sequential):
    def __init__(self, in_channels=256, size=7, fc_dims=(1024, 1024), conv_dims=(), conv_norm=""):
        """conv_dims / conv_norm: MODEL.ROI_BOX_HEAD.NUM_CONV x CONV_DIM / NORM -- ViTDet uses four 3x3 convs of 256 with
        "LN" before one FC (configs/Base-RCNN-VitDetB.yaml:7-12); conv bias is off when a norm follows."""
        super().__init__()
        cur = in_channels
        for k, d in enumerate(conv_dims):
            assert conv_norm in ("", "LN"), conv_norm
            conv = Conv2d(cur, d, kernel_size=3, padding=1, bias=not conv_norm,
                          norm=LayerNorm2d(d) if conv_norm == "LN" else None, activation=nn.ReLU())
            self.add_module("conv{}".format(k + 1), conv)
            c2_msra_fill(conv)
            cur = d
        dim = cur * size * size
        self.add_module("flatten", nn.Flatten())
        for k, fc_dim in enumerate(fc_dims):
            fc = nn.Linear(dim, fc_dim)
            self.add_module("fc{}".format(k + 1), fc)
            self.add_module("fc_relu{}".format(k + 1), nn.ReLU())
            dim = fc_dim
            c2_xavier_fill(fc)


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    boxes = Boxes(boxes.reshape(-1, 4))
    boxes.clip(image_shape)
    boxes = boxes.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


class FastRCNNOutputLayers(nn.Module):
    def __init__(self, input_size=1024, num_classes=8, test_score_thresh=0.05, test_nms_thresh=0.5,
                 test_topk_per_image=100, smooth_l1_beta=0.0):
        super().__init__()
        self.num_classes = num_classes
        self.cls_score = nn.Linear(input_size, num_classes + 1)
        self.bbox_pred = nn.Linear(input_size, num_classes * 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in [self.cls_score, self.bbox_pred]:
            nn.init.constant_(l.bias, 0)
        self.box2box_transform = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
        self.smooth_l1_beta = smooth_l1_beta
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.loss_weight = {"loss_cls": 1.0, "loss_box_reg": 1.0}

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        return self.cls_score(x), self.bbox_pred(x)

    def losses(self, predictions, proposals):
        scores, proposal_deltas = predictions
        gt_classes = cat([p.gt_classes for p in proposals], dim=0) if len(proposals) else torch.empty(0)
        _log_classification_stats(scores, gt_classes)
        if len(proposals):
            proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            gt_boxes = cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        loss_cls = cross_entropy(scores, gt_classes, reduction="mean")
        losses = {"loss_cls": loss_cls,
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        box_dim = proposal_boxes.shape[1]
        fg_inds = nonzero_tuple((gt_classes >= 0) & (gt_classes < self.num_classes))[0]
        if pred_deltas.shape[1] == box_dim:
            fg_pred_deltas = pred_deltas[fg_inds]
        else:
            fg_pred_deltas = pred_deltas.view(-1, self.num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
        loss_box_reg = _dense_box_regression_loss([proposal_boxes[fg_inds]], self.box2box_transform,
                                                  [fg_pred_deltas.unsqueeze(0)], [gt_boxes[fg_inds]], ...,
                                                  smooth_l1_beta=self.smooth_l1_beta)
        return loss_box_reg / max(gt_classes.numel(), 1.0)

    def inference(self, predictions, proposals):
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        image_shapes = [x.image_size for x in proposals]
        result_per_image = [
            fast_rcnn_inference_single_image(b, s, shp, self.test_score_thresh, self.test_nms_thresh,
                                             self.test_topk_per_image)
            for s, b, shp in zip(scores, boxes, image_shapes)
        ]
        return [x[0] for x in result_per_image], [x[1] for x in result_per_image]

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        num_prop_per_image = [len(p) for p in proposals]
        proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        predict_boxes = self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes)
        return predict_boxes.split(num_prop_per_image)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        num_inst_per_image = [len(p) for p in proposals]
        probs = F.softmax(scores, dim=-1)
        return probs.split(num_inst_per_image, dim=0)


def _log_classification_stats(pred_logits, gt_classes, prefix="fast_rcnn"):
    num_instances = gt_classes.numel()
    if num_instances == 0:
        return
    pred_classes = pred_logits.argmax(dim=1)
    bg_class_ind = pred_logits.shape[1] - 1
    fg_inds = (gt_classes >= 0) & (gt_classes < bg_class_ind)
    num_fg = fg_inds.nonzero().numel()
    fg_gt_classes = gt_classes[fg_inds]
    fg_pred_classes = pred_classes[fg_inds]
    num_false_negative = (fg_pred_classes == bg_class_ind).nonzero().numel()
    num_accurate = (pred_classes == gt_classes).nonzero().numel()
    fg_num_accurate = (fg_pred_classes == fg_gt_classes).nonzero().numel()
    storage = get_event_storage()
    storage.put_scalar(f"{prefix}/cls_accuracy", num_accurate / num_instances)
    if num_fg > 0:
        storage.put_scalar(f"{prefix}/fg_cls_accuracy", fg_num_accurate / num_fg)
        storage.put_scalar(f"{prefix}/false_negative", num_false_negative / num_fg)


def add_ground_truth_to_proposals(gt, proposals):
    assert gt is not None and len(proposals) == len(gt)
    if len(proposals) == 0:
        return proposals
    out = []
    for gt_i, proposals_i in zip(gt, proposals):
        gt_boxes = gt_i.gt_boxes if isinstance(gt_i, Instances) else gt_i
        device = proposals_i.objectness_logits.device
        gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
        gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
        gt_proposal = Instances(proposals_i.image_size, proposal_boxes=gt_boxes, objectness_logits=gt_logits)
        out.append(Instances.cat([proposals_i, gt_proposal]))
    return out


class StandardROIHeads(nn.Module):
    def __init__(self, num_classes=8, batch_size_per_image=512, positive_fraction=0.25, iou_thresholds=(0.5,),
                 iou_labels=(0, 1), proposal_append_gt=True, box_in_features=("p2", "p3", "p4", "p5"),
                 box_fc_dims=(1024, 1024), box_conv_dims=(), box_conv_norm=""):
        super().__init__()
        self.num_classes = num_classes
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.proposal_matcher = Matcher(list(iou_thresholds), list(iou_labels), allow_low_quality_matches=False)
        self.proposal_append_gt = proposal_append_gt
        self.box_in_features = box_in_features
        self.box_pooler = ROIPooler()
        self.box_head = FastRCNNConvFCHead(fc_dims=box_fc_dims, conv_dims=box_conv_dims, conv_norm=box_conv_norm)
        self.box_predictor = FastRCNNOutputLayers(input_size=box_fc_dims[-1], num_classes=num_classes)

    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
        has_gt = gt_classes.numel() > 0
        if has_gt:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes
            gt_classes[matched_labels == -1] = -1
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        sampled_fg_idxs, sampled_bg_idxs = subsample_labels(gt_classes, self.batch_size_per_image,
                                                            self.positive_fraction, self.num_classes)
        sampled_idxs = torch.cat([sampled_fg_idxs, sampled_bg_idxs], dim=0)
        return sampled_idxs, gt_classes[sampled_idxs]

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals, targets):
        if self.proposal_append_gt:
            proposals = add_ground_truth_to_proposals(targets, proposals)
        proposals_with_gt = []
        num_fg_samples, num_bg_samples = [], []
        for i, (proposals_per_image, targets_per_image) in enumerate(zip(proposals, targets)):
            SAMPLE_CTX["site"], SAMPLE_CTX["image"] = "roi", i
            has_gt = len(targets_per_image) > 0
            match_quality_matrix = pairwise_iou(targets_per_image.gt_boxes, proposals_per_image.proposal_boxes)
            matched_idxs, matched_labels = self.proposal_matcher(match_quality_matrix)
            sampled_idxs, gt_classes = self._sample_proposals(matched_idxs, matched_labels, targets_per_image.gt_classes)
            proposals_per_image = proposals_per_image[sampled_idxs]
            proposals_per_image.gt_classes = gt_classes
            if has_gt:
                sampled_targets = matched_idxs[sampled_idxs]
                for (trg_name, trg_value) in targets_per_image.get_fields().items():
                    if trg_name.startswith("gt_") and not proposals_per_image.has(trg_name):
                        proposals_per_image.set(trg_name, trg_value[sampled_targets])
            num_bg_samples.append((gt_classes == self.num_classes).sum().item())
            num_fg_samples.append(gt_classes.numel() - num_bg_samples[-1])
            proposals_with_gt.append(proposals_per_image)
        storage = get_event_storage()
        storage.put_scalar("roi_head/num_fg_samples", sum(num_fg_samples) / max(len(num_fg_samples), 1))
        storage.put_scalar("roi_head/num_bg_samples", sum(num_bg_samples) / max(len(num_bg_samples), 1))
        return proposals_with_gt

    def forward(self, images, features, proposals, targets=None):
        del images
        if self.training:
            assert targets, "'targets' argument is required during training"
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        if self.training:
            losses = self._forward_box(features, proposals)
            return proposals, losses
        pred_instances = self._forward_box(features, proposals)
        return pred_instances, {}

    def _forward_box(self, features, proposals):
        features = [features[f] for f in self.box_in_features]
        box_features = self.box_pooler(features, [x.proposal_boxes for x in proposals])
        box_features = self.box_head(box_features)
        predictions = self.box_predictor(box_features)
        del box_features
        if self.training:
            return self.box_predictor.losses(predictions, proposals)
        pred_instances, _ = self.box_predictor.inference(predictions, proposals)
        return pred_instances


# ---------------------------------------------------------------------------------------------------
# GeneralizedRCNN (detectron2/modeling/meta_arch/rcnn.py)
# ---------------------------------------------------------------------------------------------------
class GeneralizedRCNN(nn.Module):
    def __init__(self, num_classes=8, pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), freeze_at=2,
                 bottom_up=None, fpn_in_features=None, anchor_sizes=None, backbone=None, rpn_conv_dims=(-1,),
                 box_fc_dims=(1024, 1024), box_conv_dims=(), box_conv_norm="", **kwargs):
        super().__init__()
        # backbone: a complete pyramid module returning {"p2".."p6"} (oracle/vit_ref.SimpleFeaturePyramid for
        # build_vitdet_*_backbone, aldi/backbone.py:37-64), used as is instead of FPN(bottom_up)
        # bottom_up: any module exposing _out_feature_strides / _out_feature_channels (e.g. oracle/convnext_ref.ConvNeXt
        # with fpn_in_features (0, 1, 2, 3): build_convnext_fpn_backbone, aldi/backbone.py:373-392)
        if backbone is not None:
            self.backbone = backbone
        elif bottom_up is None:
            self.backbone = FPN(ResNet(freeze_at=freeze_at))
        else:
            self.backbone = FPN(bottom_up, in_features=fpn_in_features)
        self.proposal_generator = RPN(conv_dims=rpn_conv_dims)
        if anchor_sizes is not None:
            self.proposal_generator.anchor_generator = DefaultAnchorGenerator(sizes=anchor_sizes)
        self.roi_heads = StandardROIHeads(num_classes=num_classes, box_fc_dims=box_fc_dims, box_conv_dims=box_conv_dims,
                                          box_conv_norm=box_conv_norm)
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self):
        return self.pixel_mean.device

    def preprocess_image(self, batched_inputs):
        images = [x["image"].to(self.device) for x in batched_inputs]
        images = [(x - self.pixel_mean) / self.pixel_std for x in images]
        return ImageList.from_tensors(images, self.backbone.size_divisibility)

    def forward(self, batched_inputs):
        if not self.training:
            return self.inference(batched_inputs)
        images = self.preprocess_image(batched_inputs)
        if "instances" in batched_inputs[0]:
            gt_instances = [x["instances"].to(self.device) for x in batched_inputs]
        else:
            gt_instances = None
        features = self.backbone(images.tensor)
        proposals, proposal_losses = self.proposal_generator(images, features, gt_instances)
        _, detector_losses = self.roi_heads(images, features, proposals, gt_instances)
        losses = {}
        losses.update(detector_losses)
        losses.update(proposal_losses)
        return losses

    def inference(self, batched_inputs, detected_instances=None, do_postprocess=True):
        assert not self.training
        images = self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        proposals, _ = self.proposal_generator(images, features, None)
        results, _ = self.roi_heads(images, features, proposals, None)
        assert not do_postprocess, "oracle: detector_postprocess (eval path) is out of scope (SURVEY.md §2 row 10)"
        return results
