"""Faster R-CNN R50-FPN forward / backward on flat parameter buffers, every arithmetic op a C-ABI kernel call.

This is the B200 replacement of what `model(batched_inputs)` / `model.inference(...)` execute inside
Detectron2 for the reference (aldi/model.py:27-29 -> detectron2 GeneralizedRCNN):
  * parameters live in ONE flat fp32 buffer per model (student / teacher) so the EMA update and the
    optimizer step are single kernels (aldi/ema.py:32-50 loops over ~300 tensors instead);
  * activations are channels-last in the compute dtype (bf16 -> tcgen05 kernels, fp32 -> parity kernels);
  * FrozenBN / bias / ReLU / residual / FPN top-down add are conv epilogues; ReLU backward masks and
    residual joins are data-gradient epilogues; stride-2 1x1 convs, p6 and the stride-2 scatter of their
    gradients are strided views handed to TMA — no kernels;
  * backward is written out explicitly (no autograd graph): the architecture is fixed.
There is no PyTorch-op fallback: torch only allocates memory.
"""
import math
import os
from collections import OrderedDict

import torch

from . import arch, ops
from . import lib as _l

FUSED = OrderedDict([("rpn_head", ("rpn_obj", "rpn_delta")), ("predictor", ("cls_score", "bbox_pred"))])
SCALE_CLAMP = math.log(1000.0 / 16)


def _pad64(c):
    return (c + 63) // 64 * 64


class FlatLayout:
    """Detectron2 state_dict keys <-> ranges of one flat fp32 buffer (trainable range first)."""

    def __init__(self, num_classes=8, freeze_at=2, align=None, bottom_up_channels=None, head=None):
        """head: keyword arguments of arch.rcnn_specs selecting the head variant (arch.VITDET_HEADS for ViTDet)."""
        self.num_classes = num_classes
        self.align = align
        self.bottom_up_channels = bottom_up_channels
        self.head = dict(head or {})
        self.specs = arch.rcnn_specs(num_classes, freeze_at, align, bottom_up_channels, **self.head)
        member_of = {m: g for g, ms in FUSED.items() for m in ms}
        train, frozen, buffers = [], [], []
        done = set()
        for name, s in self.specs.items():
            if name in done:
                continue
            group = [name] if name not in member_of else list(FUSED[member_of[name]])
            done.update(group)
            dst = train if s.trainable else frozen
            dst.append([(g, "weight") for g in group])
            if s.bias:
                dst.append([(g, "bias") for g in group])
            if s.ln:
                dst.append([(name, "norm.weight")])
                dst.append([(name, "norm.bias")])
            if s.norm:
                for f in arch.NORM_FIELDS:
                    buffers.append([(name, "norm." + f)])
        self.entries = OrderedDict()  # (layer, field) -> (offset, numel, d2_key, d2_shape)
        off = 0
        for region, groups in (("train", train), ("frozen", frozen), ("buffers", buffers)):
            if region == "frozen":
                self.num_trainable = off
            for group in groups:
                off = (off + 3) // 4 * 4  # 16-byte alignment of every (fused) tensor
                for layer, field in group:
                    s = self.specs[layer]
                    shape = arch.d2_shape(layer, s) if field == "weight" else (s.cout,)
                    n = 1
                    for d in shape:
                        n *= d
                    self.entries[(layer, field)] = (off, n, "%s.%s" % (s.key, field), shape)
                    off += n
        self.numel = (off + 3) // 4 * 4
        self.num_trainable = (self.num_trainable + 3) // 4 * 4

    def to_internal(self, layer, field, t):
        """Detectron2 tensor -> flat-buffer element order (convs OHWI, fc1 (h,w,c) input order)."""
        if field != "weight":
            return t.reshape(-1)
        if layer == "fc1":
            return t.reshape(t.shape[0], 256, 7, 7).permute(0, 2, 3, 1).reshape(-1)
        if t.dim() == 4:
            return t.permute(0, 2, 3, 1).reshape(-1)
        return t.reshape(-1)

    def from_internal(self, layer, field, flat, shape):
        if field != "weight":
            return flat.reshape(shape).clone()
        if layer == "fc1":
            return flat.reshape(shape[0], 7, 7, 256).permute(0, 3, 1, 2).reshape(shape).contiguous()
        if len(shape) == 4:
            return flat.reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2).contiguous()
        return flat.reshape(shape).clone()

    def pack_state_dict(self, sd):
        flat = torch.zeros(max(self.numel, 4), dtype=torch.float32)
        missing = []
        for (layer, field), (off, n, key, shape) in self.entries.items():
            if key not in sd:
                missing.append(key)
                continue
            t = sd[key].detach().to("cpu", torch.float32)
            assert tuple(t.shape) == tuple(shape), (key, tuple(t.shape), shape)
            flat[off:off + n] = self.to_internal(layer, field, t)
        if missing:
            raise KeyError("missing keys in state_dict: %s" % missing[:5])
        return flat

    def unpack_state_dict(self, flat):
        flat = flat.detach().to("cpu")
        out = OrderedDict()
        for (layer, field), (off, n, key, shape) in self.entries.items():
            out[key] = self.from_internal(layer, field, flat[off:off + n], shape)
        return out


class LayerGeom:
    """Geometry of one executed GEMM layer (after head fusion)."""

    def __init__(self, name, cin, cout, k, pad, stride, members, norm, trainable, bias=True, ln=False):
        self.name, self.cin, self.cout, self.k, self.pad, self.stride = name, cin, cout, k, pad, stride
        self.members, self.norm, self.trainable = members, norm, trainable
        self.bias, self.ln = bias, ln     # has a bias vector / is followed by a trainable channel LayerNorm
        self.cin_p, self.cout_p = _pad64(cin), _pad64(cout)


class DetectorWeights:
    """One model's parameters: flat fp32 master + GEMM operands + folded FrozenBN, refreshed by kernels."""

    def __init__(self, layout, flat, dtype, bottom_up=None, split_parts=0):
        """split_parts = 2 | 3 (dtype fp32): the split-bf16 parity mode of the tensor-core path -- activations stay fp32,
        every GEMM runs on the tcgen05 kernels as 3 | 6 bf16 product terms (ops._conv_split)."""
        self.layout, self.flat, self.dtype = layout, flat, dtype
        self.dev = flat.device
        self.split_parts = int(split_parts)
        assert not self.split_parts or (dtype == torch.float32 and bottom_up is None)
        # a bottom-up that is not the ResNet-50 of this table (aldi_b200.convnext.ConvNeXtBackbone): own flat buffer,
        # forward(images, sizes, keep_masks, save) -> {0..3: stage outputs}, backward({stage: gradient})
        self.bottom_up = bottom_up
        self.geom = OrderedDict()
        sp = layout.specs
        member_of = {m: g for g, ms in FUSED.items() for m in ms}
        for name, s in sp.items():
            if name in member_of:
                g = member_of[name]
                if g in self.geom:
                    continue
                ms = FUSED[g]
                self.geom[g] = LayerGeom(g, s.cin, sum(sp[m].cout for m in ms), s.k, s.pad, s.stride, ms, False,
                                         s.trainable)
            else:
                self.geom[name] = LayerGeom(name, s.cin, s.cout, s.k, s.pad, s.stride, (name,), s.norm, s.trainable,
                                            bias=s.bias or s.norm, ln=s.ln)
        # head variants (arch.rcnn_specs): hidden 3x3 convs of the RPN head, 3x3 + LayerNorm convs of the box head
        self.rpn_convs = [n for n in self.geom if n.startswith("rpn_conv")]
        self.box_convs = [n for n in self.geom if n.startswith("box_conv")]
        self.pyramid = bottom_up is not None and getattr(bottom_up, "is_pyramid", False)
        # stem: bf16 path runs it as a GEMM over the fused normalise+im2col buffer (K = 147 -> 192)
        self.stem_gemm = dtype == torch.bfloat16 or bool(self.split_parts)
        self.fwd, self.dgrad, self.scale, self.shift = {}, {}, {}, {}
        self.fwd_split, self.dgrad_split = {}, {}
        self.scat, self.scat_split = {}, {}
        for name, g in self.geom.items():
            taps = g.k * g.k
            if name == "stem":
                kdim = 256 if self.stem_gemm else taps * 4
                self.fwd[name] = torch.zeros(g.cout_p, kdim, device=self.dev, dtype=dtype)
            else:
                self.fwd[name] = torch.zeros(g.cout_p, taps * g.cin_p, device=self.dev, dtype=dtype)
            if g.norm:
                self.scale[name] = torch.zeros(g.cout_p, device=self.dev)
            self.shift[name] = torch.zeros(g.cout_p, device=self.dev)
        self.no_dgrad = {"stem", "res3.0.conv1", "res3.0.shortcut", "fpn_lateral2", "img_align.out", "ins_align.out"}
        if bottom_up is not None:
            self.no_dgrad.discard("fpn_lateral2")   # every stage of that bottom-up trains
        self._tables = {}

    # ---- flat views -------------------------------------------------------------------------------
    def view(self, layer, field, buf=None):
        buf = self.flat if buf is None else buf
        g = self.geom[layer]
        off, _, _, _ = self.layout.entries[(g.members[0], field)]
        n = sum(self.layout.entries[(m, field)][1] for m in g.members)
        return buf[off:off + n]

    def enable_dgrad(self):
        for name, g in self.geom.items():
            if g.trainable and name not in self.no_dgrad and name not in self.dgrad:
                self.dgrad[name] = torch.zeros(g.cin_p, g.k * g.k * g.cout_p, device=self.dev, dtype=self.dtype)
        # sparse RPN backward (csrc/rpn_sparse.cu): [(tap, cin)][cout] operand that turns gathered hidden-gradient rows
        # into the 3x3 neighbourhood rows scattered back onto the FPN feature gradient
        if "rpn_conv" in self.geom:
            g = self.geom["rpn_conv"]
            self.scat["rpn_conv"] = torch.zeros(g.k * g.k * g.cin_p, g.cout_p, device=self.dev, dtype=self.dtype)
        self._tables = {}

    # ---- operand refresh: ONE launch per table (csrc/optim.cu refresh_kernel) ------------------------------------
    def _descs(self, trainable_only):
        """Descriptor list re-deriving the GEMM operands / folded FrozenBN of (a subset of) the layers."""
        dt = _l.BF16 if self.dtype == torch.bfloat16 else _l.F32
        out = []

        def desc(kind, **kw):
            d = _l.RefreshDesc()
            d.kind, d.out_dtype, d.eps = kind, dt, 1e-5
            for k, v in kw.items():
                setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
            out.append(d)

        for name, g in self.geom.items():
            taps = g.k * g.k
            w = self.view(name, "weight")
            bn = {}
            if g.norm:
                bn = dict(bn_w=self.view(name, "norm.weight"), bn_b=self.view(name, "norm.bias"),
                          bn_mean=self.view(name, "norm.running_mean"), bn_var=self.view(name, "norm.running_var"))
            if not trainable_only:
                # FrozenBN buffers / biases of frozen layers only move for the EMA teacher (aldi/ema.py covers buffers, T7)
                if g.norm:
                    desc(2, out=self.scale[name], out2=self.shift[name], cout=g.cout, **bn)
            if not g.norm and g.bias and (g.trainable or not trainable_only):
                desc(3, w=self.view(name, "bias"), out2=self.shift[name], cout=g.cout)
            if trainable_only and not g.trainable:
                continue
            if name == "stem":
                if self.stem_gemm:
                    desc(4, w=w, out=self.fwd[name], cout=g.cout, cout_p=g.cout_p)
                else:
                    desc(0, w=w, out=self.fwd[name], cout=g.cout, taps=taps, cin=3, cout_p=g.cout_p, cin_p=4)
            else:
                desc(0, w=w, out=self.fwd[name], cout=g.cout, taps=taps, cin=g.cin, cout_p=g.cout_p, cin_p=g.cin_p)
            if name in self.dgrad:
                desc(1, w=w, out=self.dgrad[name], cout=g.cout, taps=taps, cin=g.cin, cout_p=g.cout_p, cin_p=g.cin_p, **bn)
            if name in self.scat:
                assert g.cin == g.cin_p and not g.norm
                desc(5, w=w, out=self.scat[name], cout=g.cout, taps=taps, cin=g.cin, cout_p=g.cout_p, cin_p=g.cin_p)
        return out

    def _table(self, trainable_only):
        key = bool(trainable_only)
        if key not in self._tables:
            L = _l.load()
            descs = self._descs(trainable_only)
            arr = (_l.RefreshDesc * len(descs))(*descs)
            starts, tot = [], 0
            for d in descs:
                starts.append(tot)
                tot += int(L.aldi_refresh_blocks(_l.ctypes.byref(d)))
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
            st = torch.tensor(starts, dtype=torch.int32).to(self.dev)
            self._tables[key] = (raw, st, len(descs), tot)
        return self._tables[key]

    def refresh(self, trainable_only=False):
        """Re-derive operands from the master weights (after load / optimizer step / EMA update).
        trainable_only: the student after an optimizer step — frozen layers and FrozenBN buffers did not move."""
        raw, st, n, tot = self._table(trainable_only)
        ops.call("aldi_refresh_operands", raw, st, n, tot)
        if self.split_parts:
            for name, g in self.geom.items():
                if trainable_only and not g.trainable:
                    continue
                self.fwd_split[name] = ops.split_bf16(self.fwd[name], self.split_parts)
                if name in self.dgrad:
                    self.dgrad_split[name] = ops.split_bf16(self.dgrad[name], self.split_parts)
                if name in self.scat:
                    self.scat_split[name] = ops.split_bf16(self.scat[name], self.split_parts)

    def operand(self, name, dgrad=False, scatter=False):
        """The GEMM operand a conv call takes: the packed tensor, or its bf16 split in the split parity mode."""
        if self.split_parts:
            return (self.scat_split if scatter else self.dgrad_split if dgrad else self.fwd_split)[name]
        return (self.scat if scatter else self.dgrad if dgrad else self.fwd)[name]


# ---------------------------------------------------------------------------------------------------
PIXEL_MEAN = (103.530, 116.280, 123.675)
PIXEL_STD = (1.0, 1.0, 1.0)
ANCHOR_SIZES = ((32,), (64,), (128,), (256,), (512,))
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
FPN_STRIDES = (4, 8, 16, 32, 64)


def cell_anchors(anchor_sizes=None):
    """detectron2 DefaultAnchorGenerator.generate_cell_anchors (float64 math, stored as fp32)."""
    out = []
    for sizes in (anchor_sizes or ANCHOR_SIZES):
        lvl = []
        for size in sizes:
            area = size ** 2.0
            for ar in ANCHOR_RATIOS:
                w = math.sqrt(area / ar)
                h = ar * w
                lvl.append(torch.tensor([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0]).tolist())
        out.append(lvl)
    return out


class Detector:
    """Stateless executor: runs the trunk / heads of ONE DetectorWeights on a batch and, for the student,
    the explicit backward that accumulates into a flat gradient buffer."""

    RPN_CH = 16    # fp32 head output row: 3 logits + 12 deltas (+1 pad)
    PRED_CH = 64   # fp32 predictor row: K+1 logits + 4K deltas, padded

    def __init__(self, num_classes=8, anchor_sizes=None, pixel_mean=None, pixel_std=None):
        self.K = num_classes
        self.cells = cell_anchors(anchor_sizes)
        # MODEL.PIXEL_MEAN / PIXEL_STD of the cfg (detectron2 GeneralizedRCNN.preprocess_image)
        self.pixel_mean = tuple(pixel_mean) if pixel_mean is not None else PIXEL_MEAN
        self.pixel_std = tuple(pixel_std) if pixel_std is not None else PIXEL_STD

    # ---- generic layer ------------------------------------------------------------------------------
    @staticmethod
    def conv(W, name, x, *, relu=False, residual=None, res_mode=0, out=None, out_dtype=None, cout_store=None):
        g = W.geom[name]
        xv = x[:, ::2, ::2, :] if (g.stride == 2 and g.k == 1) else x
        n, h, w, _ = xv.shape
        if out is None:
            ho, wo = h + 2 * g.pad - g.k + 1, w + 2 * g.pad - g.k + 1   # stride-2 1x1 convs were turned into views
            out = torch.empty(n, ho, wo, g.cout_p, device=x.device, dtype=out_dtype or x.dtype)
        ops.conv(xv, W.operand(name), out, taps_h=g.k, taps_w=g.k, pad_h=g.pad, pad_w=g.pad, scale=W.scale.get(name),
                 bias=W.shift[name], residual=residual, res_mode=res_mode, relu=relu, cout_store=cout_store)
        return out

    # ---- trunk forward ---------------------------------------------------------------------------------
    def backbone(self, W, images_u8, sizes, save, keep_masks=None):
        """images_u8: (N,3,H,W) uint8 on device (already on the padded canvas); sizes: (N,2) int32 valid (h,w).
        Returns (features dict p2..p6 + res2..res5, saved-activation dict or None)."""
        n, _, hp, wp = images_u8.shape
        assert hp % 32 == 0 and wp % 32 == 0
        dev, dt = images_u8.device, W.dtype
        if W.pyramid:
            # ViTDet: the backbone IS the pyramid (SimpleFeaturePyramid); p6 = LastLevelMaxPool(kernel 1, stride 2) of p5
            feats = dict(W.bottom_up.forward(images_u8, sizes, keep_masks=keep_masks, save=save))
            feats["p6"] = feats["p5"][:, ::2, ::2, :]
            return feats, ({} if save else None)
        if W.bottom_up is not None:
            outs = W.bottom_up.forward(images_u8, sizes, keep_masks=keep_masks, save=save)
            feats = {"res%d" % (i + 2): outs[i] for i in range(4)}
            return self._fpn_forward(W, feats, {} if save else None, save)
        mean, std = ops.host_floats(self.pixel_mean), ops.host_floats(self.pixel_std)
        g = W.geom["stem"]
        ho, wo = hp // 2, wp // 2
        stem_out = torch.empty(n, ho, wo, 64, device=dev, dtype=dt)
        if W.stem_gemm:
            # 7x7/2 conv == 4x4/1 conv over the normalised 2x2 space-to-depth map; the 4 column taps of an output
            # pixel are 64 contiguous values, read through an overlapping-row view (row stride 16) as ONE TMA row
            wpad = wo + 4
            s2d = torch.empty(n, ho, wpad, 16, device=dev, dtype=dt)
            ops.call("aldi_stem_s2d_f32" if W.split_parts else "aldi_stem_s2d", images_u8, sizes, s2d, n, hp, wp, mean, std)
            if W.split_parts:
                s2d = ops.split_bf16(s2d, W.split_parts)     # the overlapping-row view is taken of every part
            view = s2d.as_strided((n, ho, wo, 64), (ho * wpad * 16, wpad * 16, 16, 1))
            ops.conv(view, W.operand("stem"), stem_out, taps_h=4, taps_w=1, pad_h=2, pad_w=0, scale=W.scale["stem"],
                     bias=W.shift["stem"], relu=True, algo_cin=147 / 4.0)
            del view, s2d
        else:
            x0 = torch.empty(n, hp, wp, 4, device=dev, dtype=torch.float32)
            ops.call("aldi_preprocess", images_u8, sizes, x0, n, hp, wp, hp, wp, mean, std)
            ops.conv(x0, W.fwd["stem"], stem_out, taps_h=7, taps_w=7, pad_h=3, pad_w=3, stride=2, scale=W.scale["stem"],
                     bias=W.shift["stem"], relu=True)
            del x0
        x = torch.empty(n, ho // 2, wo // 2, 64, device=dev, dtype=dt)
        ops.call("aldi_maxpool3x3s2", stem_out, x, _l.BF16 if dt == torch.bfloat16 else _l.F32, n, ho, wo, 64)
        del stem_out
        saved = {} if save else None
        feats = {}
        for si, nblk in enumerate(arch.RES_DEPTHS):
            stage = si + 2
            for b in range(nblk):
                p = "res%d.%d." % (stage, b)
                h1 = self.conv(W, p + "conv1", x, relu=True)
                h2 = self.conv(W, p + "conv2", h1, relu=True)
                sc = self.conv(W, p + "shortcut", x) if b == 0 else x
                out = self.conv(W, p + "conv3", h2, relu=True, residual=sc, res_mode=1)
                if save and W.geom[p + "conv1"].trainable:
                    saved[p] = (x, h1, h2)
                x = out
            feats["res%d" % stage] = x
        return self._fpn_forward(W, feats, saved, save)

    def _fpn_forward(self, W, feats, saved, save):
        # FPN top-down: lateral 1x1 with the nearest-2x upsampled coarser map added in the epilogue
        prev = None
        for lvl in (5, 4, 3, 2):
            r = feats["res%d" % lvl]
            prev = self.conv(W, "fpn_lateral%d" % lvl, r, residual=prev, res_mode=2 if prev is not None else 0)
            feats["p%d" % lvl] = self.conv(W, "fpn_output%d" % lvl, prev)
            if save:
                saved["prev%d" % lvl] = prev
        feats["p6"] = feats["p5"][:, ::2, ::2, :]  # LastLevelMaxPool(kernel 1, stride 2) == strided view
        return feats, saved

    # ---- RPN ----------------------------------------------------------------------------------------
    def levels(self, feats):
        shapes = [tuple(feats["p%d" % l].shape[1:3]) for l in (2, 3, 4, 5, 6)]
        return ops.make_rpn_levels(shapes, FPN_STRIDES, self.cells, self.RPN_CH, SCALE_CLAMP, 0.0)

    def rpn_head(self, W, feats, lv, save):
        n = feats["p2"].shape[0]
        dev = feats["p2"].device
        rpn_out = torch.zeros(n, lv.total_locs, self.RPN_CH, device=dev, dtype=torch.float32)
        ts = []
        for i, l in enumerate((2, 3, 4, 5, 6)):
            p = feats["p%d" % l]
            hidden = []
            t = p
            for name in W.rpn_convs:                      # one 3x3 conv + ReLU, or ViTDet's two (RPN.CONV_DIMS [-1, -1])
                t = self.conv(W, name, t, relu=True)
                hidden.append(t)
            h, w = p.shape[1], p.shape[2]
            view = rpn_out.as_strided((n, h, w, self.RPN_CH),
                                      (lv.total_locs * self.RPN_CH, w * self.RPN_CH, self.RPN_CH, 1),
                                      lv.loc_off[i] * self.RPN_CH)
            self.conv(W, "rpn_head", t, out=view, cout_store=15)
            ts.append((hidden[0] if len(hidden) == 1 else tuple(hidden)) if save else None)
        return rpn_out, ts

    def proposals(self, rpn_out, lv, sizes, pre_topk, post_topk, nms_thresh=0.7, err_flag=None):
        n = rpn_out.shape[0]
        dev = rpn_out.device
        stride = sum(min(lv.h[i] * lv.w[i] * lv.num_anchors, pre_topk) for i in range(lv.num_levels))
        cb = torch.empty(n, stride, 4, device=dev)
        cs = torch.empty(n, stride, device=dev)
        cc = torch.empty(n, stride, dtype=torch.int32, device=dev)
        ci = torch.empty(n, stride, dtype=torch.int32, device=dev)
        cv = torch.empty(n, stride, dtype=torch.uint8, device=dev)
        wsb = int(_l.load().aldi_rpn_topk_workspace_bytes(n, lv.num_levels))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        ops.call("aldi_rpn_topk_decode", rpn_out, _l.ctypes.byref(lv), n, pre_topk, sizes, cb, cs, cc, ci, cv, stride,
                 err_flag, ws, wsb)
        # batched_nms by level: every level's candidates are already score-sorted, so the levels are compacted,
        # masked and scanned in parallel and merged by rank (aldi_nms_segmented == aldi_nms_sorted on these inputs)
        lens = [min(lv.h[i] * lv.w[i] * lv.num_anchors, pre_topk) for i in range(lv.num_levels)]
        offs = [sum(lens[:i]) for i in range(lv.num_levels)]
        out = self.nms_segmented(cb, cs, cv, offs, lens, nms_thresh, post_topk)
        out["cand"] = (cb, cs, cc, ci, cv)
        return out

    @staticmethod
    def nms_segmented(cb, cs, cv, offs, lens, thresh, post_topk):
        n, stride = cs.shape
        dev = cs.device
        L = _l.load()
        k = len(lens)
        c_off, c_len = (_l.ctypes.c_int * k)(*offs), (_l.ctypes.c_int * k)(*lens)
        wsb = int(L.aldi_nms_segmented_workspace_bytes(n, k, c_len, post_topk))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        ob = torch.zeros(n, post_topk, 4, device=dev)
        osc = torch.zeros(n, post_topk, device=dev)
        oc = torch.zeros(n, post_topk, dtype=torch.int32, device=dev)
        osrc = torch.zeros(n, post_topk, dtype=torch.int32, device=dev)
        cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        ops.call("aldi_nms_segmented", cb, cs, cv, n, stride, k, c_off, c_len, thresh, post_topk, ws, wsb, ob, osc, oc, osrc, cnt)
        return {"boxes": ob, "scores": osc, "cats": oc, "src": osrc, "count": cnt}

    @staticmethod
    def nms(cb, cs, cc, cv, counts, thresh, post_topk):
        n, stride = cs.shape
        dev = cs.device
        L = _l.load()
        wsb = int(L.aldi_nms_workspace_bytes(n, stride))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        ob = torch.zeros(n, post_topk, 4, device=dev)
        osc = torch.zeros(n, post_topk, device=dev)
        oc = torch.zeros(n, post_topk, dtype=torch.int32, device=dev)
        osrc = torch.zeros(n, post_topk, dtype=torch.int32, device=dev)
        cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        ops.call("aldi_nms_sorted", cb, cs, cc, cv, counts, n, stride, thresh, post_topk, ws, wsb, ob, osc, oc, osrc, cnt)
        return {"boxes": ob, "scores": osc, "cats": oc, "src": osrc, "count": cnt}

    # ---- box head -------------------------------------------------------------------------------------
    @staticmethod
    def _roi_groups(groups, m, n):
        """groups: [(first_row, rows, first_image, images)] — `roi_batch` holds image indices LOCAL to its group (two
        passes batched into one forward keep their own RoI sampling); default: one group over everything."""
        return groups if groups is not None else [(0, m, 0, n)]

    def box_head(self, W, feats, rois, roi_batch, save, groups=None):
        m = rois.shape[0]
        dev, dt = rois.device, W.dtype
        plv = [feats["p%d" % l] for l in (2, 3, 4, 5)]
        pooled = torch.empty(m, 7, 7, 256, device=dev, dtype=dt)
        for r0, rows, i0, ni in self._roi_groups(groups, m, plv[0].shape[0]):
            ops.roi_align([p[i0:i0 + ni] for p in plv], rois[r0:r0 + rows], roi_batch[r0:r0 + rows],
                          out=pooled[r0:r0 + rows], scales=[1.0 / s for s in FPN_STRIDES[:4]])
        convs = []
        a = pooled
        dtc = _l.BF16 if dt == torch.bfloat16 else _l.F32
        for name in W.box_convs:
            # FastRCNNConvFCHead with NORM "LN": 3x3 conv (no bias) -> channel LayerNorm -> ReLU on the 7x7 RoI maps
            u = self.conv(W, name, a)
            y = torch.empty_like(u)
            stats = torch.empty(m * 49, 2, device=dev)
            ops.call("aldi_layernorm_forward", u, W.view(name, "norm.weight"), W.view(name, "norm.bias"), 1e-6, m * 49, 256, 256,
                     dtc, y, stats)
            r = torch.empty_like(y)
            ops.call("aldi_relu", y, None, r, y.numel(), dtc)
            convs.append((a, u, stats, y))
            a = r
        x = a.view(1, 1, m, 7 * 7 * 256)
        f1 = self.conv(W, "fc1", x, relu=True)
        f2 = self.conv(W, "fc2", f1, relu=True) if "fc2" in W.geom else f1
        pred = torch.zeros(1, 1, m, self.PRED_CH, device=dev, dtype=torch.float32)
        self.conv(W, "predictor", f2, out=pred, cout_store=5 * self.K + 1)
        if not save:
            return pred.view(m, self.PRED_CH), None
        return pred.view(m, self.PRED_CH), ((x, f1, f2, convs) if W.box_convs else (x, f1, f2))

    def detections(self, pred, props, sizes, score_thresh, nms_thresh=0.5, topk=100):
        """FastRCNNOutputLayers.inference on (N*P) predictions -> per-image detections (score order)."""
        n, p = props["scores"].shape
        dev = pred.device
        cstride = min(p * self.K, 16384)
        cb = torch.empty(n, cstride, 4, device=dev)
        cs = torch.empty(n, cstride, device=dev)
        cc = torch.empty(n, cstride, dtype=torch.int32, device=dev)
        csrc = torch.empty(n, cstride, dtype=torch.int32, device=dev)
        cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        ops.call("aldi_roi_inference_candidates", pred, pred.shape[1], props["boxes"], props["count"], p, n, self.K, sizes,
                 score_thresh, ops.host_floats((10.0, 10.0, 5.0, 5.0)), SCALE_CLAMP, cb, cs, cc, csrc, cnt, cstride)
        return self.nms(cb, cs, cc, None, cnt, nms_thresh, topk)

    # ---- domain-alignment discriminators (aldi/align.py:71-136; off in the shipped ALDI++ configs) -----------
    def align_forward(self, W, G, feats, f2, roi_count, roi_batch_size, cfg, labeled, gscale, loss_out):
        """`AlignMixin.forward(do_align=True)` after the detector forward: image-level ConvDiscriminator on
        feats[cfg.img_da_layer] and instance-level FCDiscriminator on the box-head output, each behind a gradient
        reversal, BCE-with-logits against the domain label (1 = source, 0 = target).  The last Linear + loss + that
        layer's gradients are one kernel; the returned dict feeds `backward(align=...)`.
        loss_out: 2 floats (loss_da_img, loss_da_ins), accumulated."""
        label = 1.0 if labeled else 0.0
        dtc = _l.BF16 if W.dtype == torch.bfloat16 else _l.F32
        dev = f2.device if f2 is not None else feats["p2"].device
        saved = {}
        if cfg.img_da_enabled:
            if len(cfg.img_da_hidden_dims) != 1:
                raise NotImplementedError("image-level discriminator: exactly one hidden conv layer is implemented "
                                          "(DOMAIN_ADAPT.ALIGN.IMG_DA_HIDDEN_DIMS default [256])")
            x = feats[cfg.img_da_layer]
            h = self.conv(W, "img_align.conv0", x, relu=True)        # 3x3 VALID conv + bias + ReLU
            n, ho, wo, cp = h.shape
            c = W.geom["img_align.conv0"].cout
            gap = torch.zeros(n, c, device=dev)
            for i in range(n):                                        # AdaptiveAvgPool2d(1)
                ops.call("aldi_colsum", h[i], dtc, 1, ho * wo, 0, cp, c, 1.0 / (ho * wo), gap[i])
            dgap = torch.empty(n, c, device=dev)
            ops.call("aldi_domain_head_loss", gap, _l.F32, n, c, c, None, 0, W.view("img_align.out", "weight"),
                     W.view("img_align.out", "bias"), label, cfg.img_da_weight, gscale, 0, dgap, None, c,
                     W.view("img_align.out", "weight", G), W.view("img_align.out", "bias", G), loss_out[0:1])
            saved["img"] = (x, h, dgap)
        if cfg.ins_da_enabled:
            if len(cfg.ins_da_hidden_dims) != 1:
                raise NotImplementedError("instance-level discriminator: exactly one hidden layer is implemented "
                                          "(DOMAIN_ADAPT.ALIGN.INS_DA_HIDDEN_DIMS default [1024])")
            hf = self.conv(W, "ins_align.fc0", f2, relu=True)        # f2: (1, 1, M, 1024) box-head output
            m, cp = hf.shape[2], hf.shape[3]
            c = W.geom["ins_align.fc0"].cout
            dh = torch.empty_like(hf)
            ndh = torch.empty_like(hf)
            if cp > c:
                dh.zero_(); ndh.zero_()
            ops.call("aldi_domain_head_loss", hf, dtc, m, c, cp, roi_count, roi_batch_size,
                     W.view("ins_align.out", "weight"), W.view("ins_align.out", "bias"), label, cfg.ins_da_weight, gscale,
                     1, dh, ndh, cp, W.view("ins_align.out", "weight", G), W.view("ins_align.out", "bias", G),
                     loss_out[1:2])
            saved["ins"] = (f2, dh, ndh)
        return saved

    def _align_img_backward(self, W, G, saved, dP_layer):
        """grad_reverse(features) -> Conv2d(valid) -> ReLU -> GAP: the conv's own gradients use dh, the features
        receive the data gradient of -dh (aldi/helpers.py:51-63 multiplies by -1)."""
        x, h, dgap = saved
        dtc = _l.BF16 if W.dtype == torch.bfloat16 else _l.F32
        n, ho, wo, cp = h.shape
        dh, ndh = torch.empty_like(h), torch.empty_like(h)
        ops.call("aldi_gap_backward", h, dgap, dtc, n, ho * wo, cp, W.geom["img_align.conv0"].cout, 1.0 / (ho * wo), dh, ndh)
        self._wgrad(W, G, "img_align.conv0", x, dh)
        self._dgrad(W, "img_align.conv0", ndh, dP_layer, accumulate=True)

    # ---- backward -------------------------------------------------------------------------------------
    def _wgrad(self, W, G, name, x, dy):
        g = W.geom[name]
        xv = x[:, ::2, ::2, :] if (g.stride == 2 and g.k == 1) else x
        # bias gradient: aldi_wgrad_tc can sum the dy tiles it already holds in shared memory (dbias=).  Measured: the
        # four extra warps' shared-memory reads slow the MMA pipeline by as much as the 21 aldi_colsum launches cost
        # (wgrad 6.83 -> 7.36 ms, colsum 0.75 -> 0 ms per step: 30.79 vs 30.77 ms), so the separate kernel stays the
        # default; ALDI_FUSED_DBIAS=1 selects the fused path.
        has_bias = g.bias and not g.norm
        fused_bias = has_bias and dy.dtype == torch.bfloat16 and os.environ.get("ALDI_FUSED_DBIAS") == "1"
        ops.wgrad(xv, dy, W.view(name, "weight", G), taps_h=g.k, taps_w=g.k, pad_h=g.pad, pad_w=g.pad,
                  scale=W.scale.get(name), cout_store=g.cout, cin_store=g.cin,
                  dbias=W.view(name, "bias", G) if fused_bias else None, split_parts=W.split_parts)
        if has_bias and not fused_bias:
            rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
            assert dy.is_contiguous()
            ops.call("aldi_colsum", dy, _l.BF16 if dy.dtype == torch.bfloat16 else _l.F32, 1, rows, 0, dy.shape[3], g.cout,
                     1.0, W.view(name, "bias", G))

    def _dgrad(self, W, name, dy, out, *, mask=None, residual=None, accumulate=False):
        g = W.geom[name]
        outv = out[:, ::2, ::2, :] if (g.stride == 2 and g.k == 1) else out
        maskv = mask[:, ::2, ::2, :] if (mask is not None and g.stride == 2 and g.k == 1) else mask
        ops.conv(dy, W.operand(name, dgrad=True), outv, taps_h=g.k, taps_w=g.k, pad_h=g.k - 1 - g.pad, pad_w=g.k - 1 - g.pad,
                 mask=maskv, residual=residual, res_mode=1 if residual is not None else 0, accumulate=accumulate,
                 cout_store=g.cin)
        return out

    def _rpn_backward_sparse(self, W, G, feats, rpn_ts, d_rpn, lv, dfeat, rows_per_image, err_flag):
        """RPN head backward on the non-zero rows of d_rpn only (csrc/rpn_sparse.cu): compact, gather, four GEMMs on the
        gathered rows, scatter-add into the fp32 feature-gradient maps `dfeat` (p2..p5; p6 rows land in p5's map)."""
        L = _l.load()
        n, dt, dev = d_rpn.shape[0], W.dtype, d_rpn.device
        dtc = _l.BF16 if dt == torch.bfloat16 else _l.F32
        gc, gh = W.geom["rpn_conv"], W.geom["rpn_head"]
        C = gc.cin_p
        cap = n * rows_per_image
        idx = torch.empty(cap, dtype=torch.int32, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        ops.call("aldi_rpn_sparse_compact", d_rpn, dtc, n, lv.total_locs, 64, gh.cout, cap, idx, count)
        dy_g = torch.empty(1, 1, cap, 64, device=dev, dtype=dt)
        t_g = torch.empty(1, 1, cap, C, device=dev, dtype=dt)
        x_g = torch.empty(1, 1, cap, 9 * C, device=dev, dtype=dt)
        p = _l.RpnSparseParams()
        p.levels = _l.ctypes.cast(_l.ctypes.pointer(lv), _l.ctypes.c_void_p)
        p.dtype, p.channels = dtc, C
        maps = list(dfeat) + [dfeat[3][:, ::2, ::2, :]]
        for i, l in enumerate((2, 3, 4, 5, 6)):
            f, t, d = feats["p%d" % l], rpn_ts[i], maps[i]
            assert t.is_contiguous() and f.stride(3) == 1 and d.stride(3) == 1 and f.shape[1:3] == d.shape[1:3]
            p.feat[i], p.feat_sn[i], p.feat_sh[i], p.feat_sw[i] = f.data_ptr(), f.stride(0), f.stride(1), f.stride(2)
            p.hidden[i] = t.data_ptr()
            p.dfeat[i], p.dfeat_sn[i], p.dfeat_sh[i], p.dfeat_sw[i] = d.data_ptr(), d.stride(0), d.stride(1), d.stride(2)
        p.drpn, p.dstride, p.idx, p.count, p.cap = d_rpn.data_ptr(), 64, idx.data_ptr(), count.data_ptr(), cap
        p.dy_g, p.t_g, p.x_g = dy_g.data_ptr(), t_g.data_ptr(), x_g.data_ptr()
        p.err_flag = err_flag.data_ptr() if err_flag is not None else None
        _l.check(ops._launch("aldi_rpn_sparse_gather", lambda: L.aldi_rpn_sparse_gather(_l.ctypes.byref(p), ops._stream())),
                 "aldi_rpn_sparse_gather")
        # 1x1 objectness / delta heads: dW += dy^T t, db += colsum(dy), dt = (W^T dy) * (t > 0)
        ops.wgrad(t_g, dy_g, W.view("rpn_head", "weight", G), cout_store=gh.cout, cin_store=gh.cin, split_parts=W.split_parts)
        ops.call("aldi_colsum", dy_g, dtc, 1, cap, 0, 64, gh.cout, 1.0, W.view("rpn_head", "bias", G))
        dt_g = torch.empty_like(t_g)
        self._dgrad(W, "rpn_head", dy_g, dt_g, mask=t_g)
        # 3x3 conv: dW[co][tap][ci] += dt^T x_g (a 1x1 layer over the 9 x C gathered columns), db += colsum(dt),
        # feature-gradient rows dx_g[k][tap][ci] = sum_co dt[k][co] w[co][tap][ci], scattered to (y + r - 1, x + s - 1)
        ops.wgrad(x_g, dt_g, W.view("rpn_conv", "weight", G), cout_store=gc.cout, cin_store=9 * gc.cin,
                  split_parts=W.split_parts)
        ops.call("aldi_colsum", dt_g, dtc, 1, cap, 0, C, gc.cout, 1.0, W.view("rpn_conv", "bias", G))
        dx_g = torch.empty_like(x_g)
        ops.conv(dt_g, W.operand("rpn_conv", scatter=True), dx_g, cout_store=9 * C)
        _l.check(ops._launch("aldi_rpn_sparse_scatter",
                             lambda: L.aldi_rpn_sparse_scatter(_l.ctypes.byref(p), ops._ptr(dx_g), ops._stream())),
                 "aldi_rpn_sparse_scatter")

    def backward(self, W, G, feats, saved, rpn_ts, d_rpn, lv, head_saved, dpred, rois, roi_batch, on_ready=None,
                 align=None, groups=None, rpn_rows_per_image=1024, err_flag=None):
        """Accumulate d(loss)/d(params) into the flat gradient buffer G.
        d_rpn: (N, total_locs, 64) activation-dtype gradient of the RPN head outputs; dpred: (M, 64).
        on_ready(tag): called when a bucket of G ("heads", "fpn", "res5", "res4", "res3") has received its last
        contribution of this backward (data_parallel.GradReducer starts that bucket's all-reduce)."""
        on_ready = on_ready or (lambda tag: None)
        dt = W.dtype
        dtc = _l.BF16 if dt == torch.bfloat16 else _l.F32
        dev = feats["p2"].device
        n = feats["p2"].shape[0]
        # ---- box head: predictor -> fc2 -> fc1 -> RoIAlign scatter
        dP = {}
        ins = (align or {}).get("ins")
        # the RPN losses touch only the sampled anchors: run the head's backward on those rows (ALDI_DENSE_RPN_BWD=1: the
        # dense convolutions over all five levels, as cuDNN does under detectron2 -- kept as the A/B and test reference)
        sparse_rpn = d_rpn is not None and os.environ.get("ALDI_DENSE_RPN_BWD") != "1" and W.bottom_up is None
        dfeat = None
        if dpred is None and ins is None:
            if sparse_rpn:
                dfeat = [torch.zeros(feats["p%d" % l].shape, device=dev, dtype=torch.float32) for l in (2, 3, 4, 5)]
            else:
                for l in (2, 3, 4, 5):
                    dP[l] = torch.zeros_like(feats["p%d" % l])
        else:
            x, f1, f2 = head_saved[:3]
            box_convs = head_saved[3] if len(head_saved) > 3 else []
            m = f2.shape[2]
            df2 = torch.empty_like(f2)
            if dpred is not None:
                dy = dpred.view(1, 1, m, dpred.shape[1])
                self._wgrad(W, G, "predictor", f2, dy)
                self._dgrad(W, "predictor", dy, df2, mask=f2)
            if ins is not None:
                # instance-level discriminator: its hidden layer learns from dh, the box head receives -dh
                _, dh, ndh = ins
                self._wgrad(W, G, "ins_align.fc0", f2, dh)
                self._dgrad(W, "ins_align.fc0", ndh, df2, mask=f2, accumulate=dpred is not None)
            if "fc2" in W.geom:
                self._wgrad(W, G, "fc2", f1, df2)
                df1 = torch.empty_like(f1)
                self._dgrad(W, "fc2", df2, df1, mask=f1)
            else:
                df1 = df2                                      # NUM_FC 1 (ViTDet): the predictor reads fc1's output
            self._wgrad(W, G, "fc1", x, df1)
            dx = torch.empty_like(x)
            self._dgrad(W, "fc1", df1, dx)
            dxv = dx.view(m, 7, 7, 256)
            for name, (a, u, stats, y) in reversed(list(zip(W.box_convs, box_convs))):
                # conv -> LayerNorm -> ReLU, backwards
                dyl = torch.empty_like(y)
                ops.call("aldi_relu", y, dxv, dyl, y.numel(), dtc)
                du = torch.empty_like(u)
                ops.call("aldi_layernorm_backward", u, W.view(name, "norm.weight"), stats, dyl, m * 49, 256, 256, dtc, du, 0,
                         W.view(name, "norm.weight", G), W.view(name, "norm.bias", G))
                self._wgrad(W, G, name, a, du)
                da = torch.empty_like(a)
                self._dgrad(W, name, du, da)
                dxv = da
                del dyl, du
            plv = [feats["p%d" % l] for l in (2, 3, 4, 5)]
            dfeat = [torch.zeros(p.shape, device=dev, dtype=torch.float32) for p in plv]
            for r0, rows, i0, ni in self._roi_groups(groups, m, n):
                ops.roi_align([p[i0:i0 + ni] for p in plv], rois[r0:r0 + rows], roi_batch[r0:r0 + rows],
                              dout=dxv[r0:r0 + rows], dfeats=[d[i0:i0 + ni] for d in dfeat],
                              scales=[1.0 / s for s in FPN_STRIDES[:4]])
            del dx, df1, df2
        if sparse_rpn:
            self._rpn_backward_sparse(W, G, feats, rpn_ts, d_rpn, lv, dfeat, rpn_rows_per_image, err_flag)
        if dfeat is not None:
            for l, d in zip((2, 3, 4, 5), dfeat):
                # first writer of dP[l]: the fp32 scatter map becomes the activation-dtype gradient (no memset + add)
                dP[l] = torch.empty_like(feats["p%d" % l])
                ops.call("aldi_cast_f32", dP[l], dtc, d, d.numel())
            del dfeat
        # ---- RPN head, dense form (weights shared over the 5 levels)
        if d_rpn is not None and not sparse_rpn:
            for i, l in enumerate((2, 3, 4, 5, 6)):
                p = feats["p%d" % l]
                h, w = p.shape[1], p.shape[2]
                hidden = rpn_ts[i] if isinstance(rpn_ts[i], tuple) else (rpn_ts[i],)
                dy = d_rpn.as_strided((n, h, w, 64), (lv.total_locs * 64, w * 64, 64, 1), lv.loc_off[i] * 64)
                self._wgrad_strided_bias(W, G, "rpn_head", hidden[-1], dy)
                dt_ = torch.empty_like(hidden[-1])
                self._dgrad(W, "rpn_head", dy, dt_, mask=hidden[-1])
                for k in reversed(range(len(W.rpn_convs))):
                    name = W.rpn_convs[k]
                    xin = hidden[k - 1] if k > 0 else p
                    self._wgrad(W, G, name, xin, dt_)
                    if k > 0:
                        dprev = torch.empty_like(xin)
                        self._dgrad(W, name, dt_, dprev, mask=xin)
                        dt_ = dprev
                    else:
                        tgt = dP[l] if l < 6 else dP[5][:, ::2, ::2, :]
                        self._dgrad(W, name, dt_, tgt, accumulate=True)
                del dt_
        if align and "img" in align:
            lname = align["img_layer"]
            tgt = dP[int(lname[1])] if lname != "p6" else dP[5][:, ::2, ::2, :]
            self._align_img_backward(W, G, align["img"], tgt)
        on_ready("heads")
        if W.pyramid:
            # ViTDet: p2..p5 are the backbone's own outputs (p6's gradient has been accumulated into p5's even positions)
            W.bottom_up.backward({"p%d" % l: dP[l] for l in (2, 3, 4, 5)})
            for tag in ("fpn", "res5", "res4", "res3"):
                on_ready(tag)
            return
        # ---- FPN: p_l = output_l(prev_l); prev_l = lateral_l(res_l) + up2(prev_{l+1})
        dprev, dres = {}, {}
        for l in (2, 3, 4, 5):
            prev = saved["prev%d" % l]
            self._wgrad(W, G, "fpn_output%d" % l, prev, dP[l])
            dprev[l] = torch.empty_like(prev)
            self._dgrad(W, "fpn_output%d" % l, dP[l], dprev[l])
            if l > 2:
                c = dprev[l]
                ops.call("aldi_sum2x2_accum", dprev[l - 1], c, dtc, c.shape[0], c.shape[1], c.shape[2], c.shape[3])
            dP[l] = None
        for l in (2, 3, 4, 5):
            r = feats["res%d" % l]
            self._wgrad(W, G, "fpn_lateral%d" % l, r, dprev[l])
            if W.bottom_up is not None:
                dres[l] = torch.empty_like(r)                 # stage outputs are LayerNorm outputs: no ReLU mask
                self._dgrad(W, "fpn_lateral%d" % l, dprev[l], dres[l])
            elif l > 2:
                dres[l] = torch.empty_like(r)
                self._dgrad(W, "fpn_lateral%d" % l, dprev[l], dres[l], mask=r)
            dprev[l] = None
        on_ready("fpn")
        if W.bottom_up is not None:
            W.bottom_up.backward({l - 2: dres[l] for l in (2, 3, 4, 5)})
            for tag in ("res5", "res4", "res3"):
                on_ready(tag)
            return
        # ---- ResNet res5 -> res3 (stem + res2 frozen: aldi configs keep D2's FREEZE_AT=2)
        for stage in (5, 4, 3):
            dout = dres[stage]
            nblk = arch.RES_DEPTHS[stage - 2]
            for b in reversed(range(nblk)):
                p = "res%d.%d." % (stage, b)
                x, h1, h2 = saved[p]
                self._wgrad(W, G, p + "conv3", h2, dout)
                dh2 = torch.empty_like(h2)
                self._dgrad(W, p + "conv3", dout, dh2, mask=h2)
                self._wgrad(W, G, p + "conv2", h1, dh2)
                dh1 = torch.empty_like(h1)
                self._dgrad(W, p + "conv2", dh2, dh1, mask=h1)
                del dh2
                self._wgrad(W, G, p + "conv1", x, dh1)
                if b == 0:
                    self._wgrad(W, G, p + "shortcut", x, dout)
                    if stage > 3:
                        # x is the previous stage's output: its gradient buffer already holds the FPN-lateral
                        # term; the two stride-2 branches scatter-accumulate into the even positions
                        dx = dres[stage - 1]
                        self._dgrad(W, p + "conv1", dh1, dx, mask=x, accumulate=True)
                        self._dgrad(W, p + "shortcut", dout, dx, mask=x, accumulate=True)
                else:
                    dx = torch.empty_like(x)
                    self._dgrad(W, p + "conv1", dh1, dx, mask=x, residual=dout)
                    dout = dx
                del dh1
            dres[stage] = None
            on_ready("res%d" % stage)

    def _wgrad_strided_bias(self, W, G, name, x, dy):
        """wgrad + bias grad where dy is a strided level view of the concatenated RPN gradient map."""
        g = W.geom[name]
        fused_bias = dy.dtype == torch.bfloat16 and os.environ.get("ALDI_FUSED_DBIAS") == "1"
        ops.wgrad(x, dy, W.view(name, "weight", G), taps_h=g.k, taps_w=g.k, pad_h=g.pad, pad_w=g.pad,
                  cout_store=g.cout, cin_store=g.cin, dbias=W.view(name, "bias", G) if fused_bias else None,
                  split_parts=W.split_parts)
        if fused_bias:
            return
        n, h, w, c = dy.shape
        dtc = _l.BF16 if dy.dtype == torch.bfloat16 else _l.F32
        # rows of one image's level slab are contiguous; images are total_locs * c apart
        ops.call("aldi_colsum", dy, dtc, n, h * w, dy.stride(0), c, g.cout, 1.0, W.view(name, "bias", G))
