"""ViTDet backbone on B200: plain ViT + SimpleFeaturePyramid, forward and explicit backward over a flat parameter buffer
(BASELINE configs[2]).

Mirror of what `build_vitdet_b_backbone` / `build_vitdet_l_backbone` instantiate (aldi/backbone.py:37-64: Detectron2's
`SimpleFeaturePyramid(net=ViT(...))` from `common/models/mask_rcnn_vitdet.py` with `square_pad = 0`, forward replaced by
`checkpointed_vit_forward`, aldi/backbone.py:21-35) with Detectron2's state_dict keys (`net.pos_embed`,
`net.patch_embed.proj`, `net.blocks.{i}.{norm1,attn.qkv,attn.proj,attn.rel_pos_h,attn.rel_pos_w,norm2,mlp.fc1,mlp.fc2}`,
`simfp_{2..5}.{k}[.norm]`), so released ViTDet checkpoints load.  Activation checkpointing (`VIT.USE_ACT_CHECKPOINT`) changes
memory, not values: every block's activations are kept (180 GB of HBM3e).

Layout: activations channels-last (N, H, W, C padded to 64).  Every Linear / 1x1 / 3x3 / transposed conv is a GEMM on the
tcgen05 implicit-GEMM kernels (`aldi_conv_tc` / `aldi_wgrad_tc`; fp32 CUDA-core kernels in parity mode):
  * patch embedding = `aldi_patchify_image` (16 x 16 x 3 rows straight from the uint8 canvas) + one GEMM;
  * attention = `aldi_attention_forward / _backward` (csrc/attn_tc.cu: QK^T, PV and the four backward products as UMMA
    tiles over 16 x 8-token TMA patches of the qkv tensor as the qkv Linear wrote it; csrc/vit.cu in parity mode); the
    decomposed relative-position term enters as ONE extra GEMM per block, q x [Rh; Rw]^T, re-laid out key-major
    (`aldi_relpos_transpose`) so the attention kernels read it per (key row / key column, query) with the query index
    fastest -- the tables' gradients and the term's share of dq are plain GEMMs too;
  * windowed blocks run on the partitioned tensor (`aldi_window_partition`, padding tokens written as zeros AFTER norm1 and
    never masked, as detectron2 does: they are keys of their window and carry qkv = bias);
  * ConvTranspose2d(k=2, s=2) = GEMM to (dy, dx, cout) columns + depth-to-space (`aldi_space_to_depth`, inverse);
  * DropPath (timm, per sample, two draws per block) rides in the residual kernel (`aldi_layerscale_forward`, gamma = 1).
Parity: tests/test_gpu_vit.py holds every stage against oracle/vit_ref.py (parity unpinned: Detectron2's vit.py is not under
/root/reference).
"""
import math
import os
from collections import OrderedDict

import torch

from . import lib as _l
from . import ops

WINDOW = 14


def _pad64(c):
    return (c + 63) // 64 * 64


def vit_config(size="b"):
    """aldi/backbone.py:37-64 on top of mask_rcnn_vitdet.py's ViT-B defaults."""
    if size == "b":
        return dict(embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.1, window_block_indexes=(0, 1, 3, 4, 6, 7, 9, 10),
                    lr_decay_rate=0.7)
    if size == "l":
        wb = tuple(list(range(0, 5)) + list(range(6, 11)) + list(range(12, 17)) + list(range(18, 23)))
        # layer-wise lr decay is only switched on for ViT-B (aldi/trainer.py:206)
        return dict(embed_dim=1024, depth=24, num_heads=16, drop_path_rate=0.4, window_block_indexes=wb, lr_decay_rate=1.0)
    raise ValueError("ViT size %r: 'b' | 'l'" % (size,))


class ViTLayout:
    """state_dict keys (relative to `backbone.`) <-> ranges of one flat fp32 buffer, plus the per-range AdamW settings
    detectron2's get_default_optimizer_params derives (aldi/backbone.py:66-84): layer-wise lr decay, no weight decay on
    torch.nn.LayerNorm parameters and on pos_embed."""

    def __init__(self, embed_dim=768, depth=12, num_heads=12, window_block_indexes=(), img_size=1024, patch_size=16,
                 pretrain_img_size=224, mlp_ratio=4.0, out_channels=256, window_size=WINDOW, lr_decay_rate=1.0, **_):
        C, hd = embed_dim, embed_dim // num_heads
        assert hd == 64, "the attention kernels are built for head dim 64 (ViT-B / ViT-L / ViT-H)"
        self.embed_dim, self.depth, self.num_heads, self.patch = C, depth, num_heads, patch_size
        self.window_blocks = set(window_block_indexes)
        self.window_size, self.out_channels = window_size, out_channels
        self.grid = img_size // patch_size
        self.pos_side = pretrain_img_size // patch_size
        hidden = int(C * mlp_ratio)
        e = OrderedDict()
        self.convt = set()
        opt = {}

        def add(key, shape, layer_id, no_wd=False):
            e[key] = tuple(shape)
            opt[key] = (lr_decay_rate ** (depth + 1 - layer_id), no_wd)

        add("net.pos_embed", (1, self.pos_side ** 2 + 1, C), 0, True)
        add("net.patch_embed.proj.weight", (C, 3, patch_size, patch_size), 0)
        add("net.patch_embed.proj.bias", (C,), 0)
        for i in range(depth):
            p, lid = "net.blocks.%d." % i, i + 1
            s = window_size if i in self.window_blocks else self.grid
            for k in ("norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"):
                add(p + k, (C,), lid, True)
            add(p + "attn.rel_pos_h", (2 * s - 1, hd), lid)          # adjacent: [Rh; Rw] is one GEMM operand
            add(p + "attn.rel_pos_w", (2 * s - 1, hd), lid)
            add(p + "attn.qkv.weight", (3 * C, C), lid); add(p + "attn.qkv.bias", (3 * C,), lid)
            add(p + "attn.proj.weight", (C, C), lid); add(p + "attn.proj.bias", (C,), lid)
            add(p + "mlp.fc1.weight", (hidden, C), lid); add(p + "mlp.fc1.bias", (hidden,), lid)
            add(p + "mlp.fc2.weight", (C, hidden), lid); add(p + "mlp.fc2.bias", (C,), lid)
        top = depth + 1
        oc = out_channels

        def conv_norm(prefix, cin, k):
            add(prefix + ".weight", (oc, cin, k, k), top)
            add(prefix + ".norm.weight", (oc,), top); add(prefix + ".norm.bias", (oc,), top)

        def convt(prefix, cin, cout):
            add(prefix + ".weight", (cin, cout, 2, 2), top); add(prefix + ".bias", (cout,), top)
            self.convt.add(prefix + ".weight")

        convt("simfp_2.0", C, C // 2)
        add("simfp_2.1.weight", (C // 2,), top); add("simfp_2.1.bias", (C // 2,), top)
        convt("simfp_2.3", C // 2, C // 4)
        conv_norm("simfp_2.4", C // 4, 1); conv_norm("simfp_2.5", oc, 3)
        convt("simfp_3.0", C, C // 2)
        conv_norm("simfp_3.1", C // 2, 1); conv_norm("simfp_3.2", oc, 3)
        conv_norm("simfp_4.0", C, 1); conv_norm("simfp_4.1", oc, 3)
        conv_norm("simfp_5.1", C, 1); conv_norm("simfp_5.2", oc, 3)
        self.entries = OrderedDict()
        off = 0
        for k, shape in e.items():
            n = 1
            for s in shape:
                n *= s
            self.entries[k] = (off, n, shape)
            off = (off + n + 3) // 4 * 4
        self.numel = off
        self.opt = opt

    def to_internal(self, key, t):
        if key in self.convt:                                  # (cin, cout, dy, dx) -> GEMM rows (dy, dx, cout) x cin
            return t.permute(2, 3, 1, 0).reshape(-1)
        if t.dim() == 4:                                       # OIHW -> OHWI
            return t.permute(0, 2, 3, 1).reshape(-1)
        return t.reshape(-1)

    def from_internal(self, key, flat, shape):
        if key in self.convt:
            return flat.reshape(shape[2], shape[3], shape[1], shape[0]).permute(3, 2, 0, 1).contiguous()
        if len(shape) == 4:
            return flat.reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2).contiguous()
        return flat.reshape(shape).clone()

    def pack(self, sd):
        flat = torch.zeros(self.numel, dtype=torch.float32)
        missing = [k for k in self.entries if k not in sd]
        if missing:
            raise KeyError("missing keys in state_dict: %s" % missing[:5])
        for k, (off, n, shape) in self.entries.items():
            t = sd[k].detach().to("cpu", torch.float32)
            assert tuple(t.shape) == shape, (k, tuple(t.shape), shape)
            flat[off:off + n] = self.to_internal(k, t)
        return flat

    def unpack(self, flat):
        flat = flat.detach().cpu()
        return OrderedDict((k, self.from_internal(k, flat[off:off + n], shape)) for k, (off, n, shape) in self.entries.items())

    def opt_segments(self):
        """[(offset, numel, lr_factor, no_weight_decay)] -- adjacent tensors with equal settings merged."""
        segs = []
        for k, (off, n, _) in self.entries.items():
            f, nowd = self.opt[k]
            end = (off + n + 3) // 4 * 4
            if segs and segs[-1][2] == f and segs[-1][3] == nowd and segs[-1][0] + segs[-1][1] == off:
                segs[-1] = (segs[-1][0], end - segs[-1][0], f, nowd)
            else:
                segs.append((off, end - off, f, nowd))
        return segs


def synthetic_state_dict(layout, seed=0, rel_pos_std=0.0):
    """Detectron2's own initialisation (vit.py `_init_weights`: trunc_normal(0.02) Linear weights, zero biases, unit
    LayerNorms, trunc_normal(0.02) pos_embed, zero rel-pos tables; SimpleFeaturePyramid: c2_msra_fill convs, default
    ConvTranspose2d init) stands in for the MAE / COCO checkpoint (`models/model_final_61ccd1.pkl`,
    configs/Base-RCNN-VitDetB.yaml:5) that cannot be fetched here.  rel_pos_std > 0 draws non-zero tables (tests)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, (_, _, shape) in layout.entries.items():
        if "rel_pos" in k:
            sd[k] = torch.randn(shape, generator=g) * rel_pos_std
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shape)
        elif len(shape) == 1:
            sd[k] = torch.ones(shape)
        elif k.startswith("simfp") and len(shape) == 4:
            fan = shape[0] * shape[2] * shape[3] if k not in layout.convt else shape[0] * 4
            sd[k] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan)
        else:
            sd[k] = torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=g)
    return sd


class _Gemm:
    """One dense / conv layer on the implicit-GEMM kernels.  `wview()` returns the fp32 master (cout, taps, cin) range of the
    flat buffer, `gview()` the same range of the gradient buffer; bias (optional) likewise by key."""

    def __init__(self, net, name, cout, cin, k=1, bias_key=None, bias_rep=1, dgrad=True, wkey=None, wview=None, gview=None):
        self.net, self.name, self.cout, self.cin, self.k = net, name, cout, cin, k
        self.cout_p, self.cin_p, self.taps = _pad64(cout), _pad64(cin), k * k
        self.bias_key, self.bias_rep = bias_key, bias_rep
        self.wview = wview or (lambda: net.view(wkey))
        self.gview = gview or (lambda: net.view(wkey, net.grad))
        dev, dt = net.flat.device, net.dtype
        self.fwd = torch.zeros(self.cout_p, self.taps * self.cin_p, device=dev, dtype=dt)
        self.bwd = torch.zeros(self.cin_p, self.taps * self.cout_p, device=dev, dtype=dt) if dgrad else None
        self.bias = torch.zeros(self.cout_p, device=dev) if bias_key else None

    def descs(self, desc):
        geo = dict(cout=self.cout, taps=self.taps, cin=self.cin, cout_p=self.cout_p, cin_p=self.cin_p)
        w = self.wview()
        desc(0, w=w, out=self.fwd, **geo)
        if self.bwd is not None:
            desc(1, w=w, out=self.bwd, **geo)
        if self.bias_key and self.bias_rep == 1:
            desc(3, w=self.net.view(self.bias_key), out2=self.bias, cout=self.cout)

    def refresh_bias(self):
        if self.bias_key and self.bias_rep > 1:            # ConvTranspose2d: one bias per output channel, (dy, dx) columns
            b = self.net.view(self.bias_key)
            co = b.numel()
            for r in range(self.bias_rep):
                self.bias[r * co:(r + 1) * co].copy_(b)

    def forward(self, x, out=None, out_dtype=None):
        n, h, w, _ = x.shape
        pad = self.k // 2
        if out is None:
            out = torch.empty(n, h, w, self.cout_p, device=x.device, dtype=out_dtype or x.dtype)
        ops.conv(x, self.fwd, out, taps_h=self.k, taps_w=self.k, pad_h=pad, pad_w=pad, bias=self.bias)
        return out

    def backward(self, x, dy, want_dx=True, dx=None, accumulate=False):
        net = self.net
        pad = self.k // 2
        # bf16: a plain bias gradient rides on the weight-gradient pass (column sums of the dy tiles it already stages in
        # shared memory, as in the ConvNeXt blocks); ALDI_VIT_COLSUM=1 keeps the separate aldi_colsum pass (A/B knob)
        fused = bool(self.bias_key) and self.bias_rep == 1 and dy.dtype == torch.bfloat16 and os.environ.get("ALDI_VIT_COLSUM") != "1"
        ops.wgrad(x, dy, self.gview(), taps_h=self.k, taps_w=self.k, pad_h=pad, pad_w=pad, cout_store=self.cout, cin_store=self.cin,
                  dbias=net.view(self.bias_key, net.grad) if fused else None)
        if self.bias_key and not fused:
            db = net.view(self.bias_key, net.grad)
            rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
            assert dy.is_contiguous()
            if self.bias_rep > 1:
                co = db.numel()
                assert dy.shape[3] == self.bias_rep * co, "transposed-conv bias gradient needs unpadded (dy, dx, cout) columns"
                ops.call("aldi_colsum", dy, net.dtc, 1, rows * self.bias_rep, 0, co, co, 1.0, db)
            else:
                for c0 in range(0, self.cout, 2048):           # aldi_colsum handles up to 2048 channels per launch
                    ops.call("aldi_colsum", dy.view(rows, dy.shape[3])[:, c0:], net.dtc, 1, rows, 0, dy.shape[3],
                             min(2048, self.cout - c0), 1.0, db[c0:])
        if not want_dx:
            return None
        if dx is None:
            dx = torch.empty(dy.shape[0], dy.shape[1], dy.shape[2], self.cin_p, device=dy.device, dtype=dy.dtype)
        ops.conv(dy, self.bwd, dx, taps_h=self.k, taps_w=self.k, pad_h=self.k - 1 - pad, pad_w=self.k - 1 - pad,
                 accumulate=accumulate, cout_store=self.cin_p if not accumulate else self.cin)
        return dx


class ViTDetBackbone:
    """`backbone` of the ViTDet detector: forward(images, sizes, keep_masks, save) -> {"p2".."p5"} channels-last maps,
    backward({"p2".."p5": gradients}).  `is_pyramid`: the detector uses the maps as they are (no FPN on top)."""

    is_pyramid = True
    prefix = "backbone."          # state_dict keys: backbone.net.*, backbone.simfp_*
    keyed_layout = True

    def __init__(self, state_dict, size="b", dtype="bf16", device="cuda:0", pixel_mean=(123.675, 116.28, 103.53),
                 pixel_std=(58.395, 57.12, 57.375), img_size=1024, **overrides):
        _l.load()   # fail loudly if the CUDA library is missing: there is no PyTorch fallback
        cfg = dict(vit_config(size) if size else {})
        cfg.update(overrides)
        self.cfg = cfg
        self.layout = ViTLayout(img_size=img_size, **cfg)
        L = self.layout
        self.C, self.depth, self.heads = L.embed_dim, L.depth, L.num_heads
        self.dtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.dtc = _l.BF16 if dtype == "bf16" else _l.F32
        self.device = torch.device(device)
        self.flat = L.pack(state_dict).to(self.device)
        self.grad = torch.zeros_like(self.flat)
        self.mean, self.std = tuple(pixel_mean), tuple(pixel_std)
        dpr = [x.item() for x in torch.linspace(0, cfg.get("drop_path_rate", 0.0), self.depth)]
        self.drop_rates = [r for r in dpr for _ in range(2)]        # two draws per block, forward order (timm DropPath)
        self.scale = 64 ** -0.5
        C, oc = self.C, L.out_channels
        assert C % 64 == 0
        g = OrderedDict()
        g["patch"] = _Gemm(self, "patch", C, 3 * L.patch ** 2, bias_key="net.patch_embed.proj.bias", dgrad=False,
                           wkey="net.patch_embed.proj.weight")
        hidden = L.entries["net.blocks.0.mlp.fc1.bias"][1]
        for i in range(self.depth):
            p = "net.blocks.%d." % i
            g[p + "qkv"] = _Gemm(self, p + "qkv", 3 * C, C, bias_key=p + "attn.qkv.bias", wkey=p + "attn.qkv.weight")
            g[p + "proj"] = _Gemm(self, p + "proj", C, C, bias_key=p + "attn.proj.bias", wkey=p + "attn.proj.weight")
            g[p + "fc1"] = _Gemm(self, p + "fc1", hidden, C, bias_key=p + "mlp.fc1.bias", wkey=p + "mlp.fc1.weight")
            g[p + "fc2"] = _Gemm(self, p + "fc2", C, hidden, bias_key=p + "mlp.fc2.bias", wkey=p + "mlp.fc2.weight")

        def convt(prefix, cin, cout):
            g[prefix] = _Gemm(self, prefix, 4 * cout, cin, bias_key=prefix + ".bias", bias_rep=4, wkey=prefix + ".weight")

        convt("simfp_2.0", C, C // 2); convt("simfp_2.3", C // 2, C // 4); convt("simfp_3.0", C, C // 2)
        for prefix, cin, k in (("simfp_2.4", C // 4, 1), ("simfp_2.5", oc, 3), ("simfp_3.1", C // 2, 1), ("simfp_3.2", oc, 3),
                               ("simfp_4.0", C, 1), ("simfp_4.1", oc, 3), ("simfp_5.1", C, 1), ("simfp_5.2", oc, 3)):
            g[prefix] = _Gemm(self, prefix, oc, cin, k=k, wkey=prefix + ".weight")
        self.gemm = g
        self._rel = {}        # (block, gh, gw) -> _Gemm over [Rh; Rw] (+ resampled tables when the grid is not the table's)
        self._pos = {}        # (gh, gw) -> fp32 (gh * gw, C) absolute position embedding
        self._pos_g = {}      # ... and the scratch its gradient is summed into
        self._table = None
        self.saved = None
        self.refresh()

    # ---- flat views ---------------------------------------------------------------------------------------------
    def view(self, key, buf=None):
        off, n, _ = self.layout.entries[key]
        return (self.flat if buf is None else buf)[off:off + n]

    def state_dict(self):
        return self.layout.unpack(self.flat)

    def opt_segments(self):
        return self.layout.opt_segments()

    def draw_keep_masks(self, n, generator=None):
        return [None if r <= 0 else (torch.rand(n, generator=generator) < (1 - r)).float() / (1 - r) for r in self.drop_rates]

    # ---- derived operands ---------------------------------------------------------------------------------------
    def refresh(self):
        """Re-derive every GEMM operand from the master weights (after load / optimizer step / EMA update): ONE launch of
        the batched refresh kernel (csrc/optim.cu) + the resampled tables of the grids seen so far."""
        layers = list(self.gemm.values()) + [r["gemm"] for r in self._rel.values()]
        if self._table is None or self._table[4] != len(layers):
            L = _l.load()
            descs = []

            def desc(kind, **kw):
                d = _l.RefreshDesc()
                d.kind, d.out_dtype, d.eps = kind, self.dtc, 1e-6
                for k, v in kw.items():
                    setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
                descs.append(d)

            for lay in layers:
                lay.descs(desc)
            arr = (_l.RefreshDesc * len(descs))(*descs)
            starts, tot = [], 0
            for d in descs:
                starts.append(tot)
                tot += int(L.aldi_refresh_blocks(_l.ctypes.byref(d)))
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            st = torch.tensor(starts, dtype=torch.int32).to(self.device)
            self._table = (raw, st, len(descs), tot, len(layers))
        for r in self._rel.values():
            self._resample_rel(r)
        raw, st, n, tot, _ = self._table
        ops.call("aldi_refresh_operands", raw, st, n, tot)
        for lay in self.gemm.values():
            lay.refresh_bias()
        for (gh, gw), pos in self._pos.items():
            self._resample_pos(pos, gh, gw)

    def _resample_pos(self, pos, gh, gw):
        L = self.layout
        src = self.view("net.pos_embed")[self.C:]              # the class-token slot is dropped (get_abs_pos)
        ops.call("aldi_bicubic_resize", src, L.pos_side, L.pos_side, pos, gh, gw, self.C, self.C, 0)

    def _pos_embed(self, gh, gw):
        if (gh, gw) not in self._pos:
            pos = torch.empty(gh * gw, self.C, device=self.device)
            self._resample_pos(pos, gh, gw)
            self._pos[(gh, gw)] = pos
        return self._pos[(gh, gw)]

    def _resample_rel(self, r):
        if r["eff"] is None:
            return
        i, gh, gw, s = r["block"], r["gh"], r["gw"], r["s"]
        p = "net.blocks.%d.attn." % i
        ops.call("aldi_linear_resize_rows", self.view(p + "rel_pos_h"), 2 * s - 1, r["eff"][:(2 * gh - 1) * 64], 2 * gh - 1, 64, 0)
        ops.call("aldi_linear_resize_rows", self.view(p + "rel_pos_w"), 2 * s - 1, r["eff"][(2 * gh - 1) * 64:], 2 * gw - 1, 64, 0)

    def _rel_gemm(self, i, gh, gw):
        """[Rh (2gh-1 rows); Rw (2gw-1 rows)] x 64 as a GEMM layer: the parameter range itself when the token grid matches
        the tables (rel_pos_h / rel_pos_w are adjacent in the flat buffer), else linearly resampled copies (get_rel_pos)."""
        key = (i, gh, gw)
        if key not in self._rel:
            s = self.layout.window_size if i in self.layout.window_blocks else self.layout.grid
            p = "net.blocks.%d.attn." % i
            nr = 2 * gh - 1 + 2 * gw - 1
            off_h = self.layout.entries[p + "rel_pos_h"][0]
            r = dict(block=i, gh=gh, gw=gw, s=s, eff=None, geff=None)
            if gh == s and gw == s:
                wview = lambda: self.flat[off_h:off_h + nr * 64]                 # noqa: E731
                gview = lambda: self.grad[off_h:off_h + nr * 64]                 # noqa: E731
            else:
                r["eff"] = torch.empty(nr * 64, device=self.device)
                r["geff"] = torch.zeros(nr * 64, device=self.device)
                wview = lambda: r["eff"]                                         # noqa: E731
                gview = lambda: r["geff"]                                        # noqa: E731
                self._resample_rel(r)
            r["gemm"] = _Gemm(self, p + "rel", nr, 64, wview=wview, gview=gview)
            # a layer added after the first refresh: derive its operands now, and rebuild the batched table next time
            ops.pack_weight(wview(), r["gemm"].fwd, cout=nr, taps=1, cin=64, cout_p=r["gemm"].cout_p, cin_p=64)
            ops.pack_weight(wview(), r["gemm"].bwd, dgrad=True, cout=nr, taps=1, cin=64, cout_p=r["gemm"].cout_p, cin_p=64)
            self._rel[key] = r
        return self._rel[key]

    # ---- pieces -------------------------------------------------------------------------------------------------
    def _ln(self, x, key, c):
        n, h, w, cp = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(n * h * w, 2, device=x.device)
        ops.call("aldi_layernorm_forward", x, self.view(key + ".weight"), self.view(key + ".bias"), 1e-6, n * h * w, c, cp,
                 self.dtc, y, stats)
        return y, stats

    def _ln_bwd(self, x, stats, dy, key, c, dx=None):
        n, h, w, cp = x.shape
        acc = dx is not None
        if dx is None:
            dx = torch.empty_like(x)
        ops.call("aldi_layernorm_backward", x, self.view(key + ".weight"), stats, dy, n * h * w, c, cp, self.dtc, dx, int(acc),
                 self.view(key + ".weight", self.grad), self.view(key + ".bias", self.grad))
        return dx

    def _attn_params(self, qkv, rel_hw, out, lse):
        b, gh, gw, _ = qkv.shape
        p = _l.AttnParams()
        p.qkv, p.batch, p.gh, p.gw, p.heads = qkv.data_ptr(), b, gh, gw, self.heads
        p.row_stride, p.batch_stride = qkv.stride(2), qkv.stride(0)
        p.rel_h, p.rel_w = rel_hw[0].data_ptr(), rel_hw[1].data_ptr()
        p.scale, p.dtype = self.scale, self.dtc
        p.out, p.out_stride, p.out_batch_stride, p.lse = out.data_ptr(), out.stride(2), out.stride(0), lse.data_ptr()
        p.impl = 0
        return p

    def _attention(self, i, xin, save):
        """qkv Linear -> relative-position products -> attention -> proj Linear on a (B, gh, gw, C) token tensor."""
        p = "net.blocks.%d." % i
        b, gh, gw, _ = xin.shape
        t = gh * gw
        qkv = self.gemm[p + "qkv"].forward(xin)
        rg = self._rel_gemm(i, gh, gw)["gemm"]
        q4 = qkv.as_strided((b, t, self.heads, 64), (qkv.stride(0), qkv.stride(2), 64, 1), qkv.storage_offset())
        rel = rg.forward(q4, out_dtype=torch.float32)          # (b, tokens, heads, pad64(2gh-1 + 2gw-1)) fp32
        # key-major copies the attention kernels read with the query index fastest (coalesced across a warp's 32 rows)
        rel_h = torch.empty(b, self.heads, gh, t, device=xin.device)
        rel_w = torch.empty(b, self.heads, gw, t, device=xin.device)
        ops.call("aldi_relpos_transpose", rel, rel.shape[3], rel_h, rel_w, b, gh, gw, self.heads, 0)
        del rel
        out = torch.empty(b, gh, gw, self.C, device=xin.device, dtype=xin.dtype)
        lse = torch.empty(b, self.heads, t, device=xin.device)
        ap = self._attn_params(qkv, (rel_h, rel_w), out, lse)
        L = _l.load()
        _l.check(ops._launch("aldi_attention_forward", lambda: L.aldi_attention_forward(_l.ctypes.byref(ap), ops._stream()),
                             4.0 * b * self.heads * t ** 2 * 64), "aldi_attention_forward")
        y = self.gemm[p + "proj"].forward(out)
        return y, ((xin, qkv, rel_h, rel_w, out, lse) if save else None)

    def _attention_bwd(self, i, saved, dy):
        p = "net.blocks.%d." % i
        xin, qkv, rel_h, rel_w, out, lse = saved
        b, gh, gw, _ = xin.shape
        t = gh * gw
        dout = self.gemm[p + "proj"].backward(out, dy)
        dqkv = torch.empty_like(qkv)
        drel_h, drel_w = torch.empty_like(rel_h), torch.empty_like(rel_w)
        delta = torch.empty_like(lse)
        ap = self._attn_params(qkv, (rel_h, rel_w), out, lse)
        ap.dout, ap.dqkv, ap.drel_h, ap.drel_w, ap.delta = (dout.data_ptr(), dqkv.data_ptr(), drel_h.data_ptr(), drel_w.data_ptr(),
                                                           delta.data_ptr())
        L = _l.load()
        _l.check(ops._launch("aldi_attention_backward", lambda: L.aldi_attention_backward(_l.ctypes.byref(ap), ops._stream()),
                             10.0 * b * self.heads * t ** 2 * 64), "aldi_attention_backward")
        # the relative-position products are a GEMM of the q view: table gradients = wgrad, their share of dq = dgrad
        r = self._rel_gemm(i, gh, gw)
        rg = r["gemm"]
        drel = torch.empty(b, t, self.heads, rg.cout_p, device=xin.device)
        ops.call("aldi_relpos_transpose", drel, rg.cout_p, drel_h, drel_w, b, gh, gw, self.heads, 1)
        if self.dtype == torch.bfloat16:
            drel_a = torch.empty(drel.shape, device=drel.device, dtype=self.dtype)
            ops.call("aldi_cast_f32", drel_a, self.dtc, drel, drel.numel())
        else:
            drel_a = drel
        q4 = qkv.as_strided((b, t, self.heads, 64), (qkv.stride(0), qkv.stride(2), 64, 1), qkv.storage_offset())
        dq4 = dqkv.as_strided((b, t, self.heads, 64), (dqkv.stride(0), dqkv.stride(2), 64, 1), dqkv.storage_offset())
        rg.backward(q4, drel_a, dx=dq4, accumulate=True)
        if r["eff"] is not None:
            # resampled tables: fold their gradient back onto the parameters (transpose of the linear interpolation)
            s = r["s"]
            ops.call("aldi_linear_resize_rows", self.view(p + "attn.rel_pos_h", self.grad), 2 * s - 1, r["geff"][:(2 * gh - 1) * 64],
                     2 * gh - 1, 64, 1)
            ops.call("aldi_linear_resize_rows", self.view(p + "attn.rel_pos_w", self.grad), 2 * s - 1, r["geff"][(2 * gh - 1) * 64:],
                     2 * gw - 1, 64, 1)
            r["geff"].zero_()
        return self.gemm[p + "qkv"].backward(xin, dqkv)

    def _residual(self, u, x, keep):
        n, h, w, cp = x.shape
        out = torch.empty_like(x)
        ops.call("aldi_layerscale_forward", u, x, None, keep, n * h * w, h * w, self.C, cp, self.dtc, out)
        return out

    def _residual_bwd(self, dy, keep):
        if keep is None:
            return dy
        n, h, w, cp = dy.shape
        du = torch.empty_like(dy)
        ops.call("aldi_layerscale_backward", dy, dy, None, keep, n * h * w, h * w, self.C, cp, self.dtc, du, None)
        return du

    def _convt(self, name, x, cout):
        """ConvTranspose2d(kernel 2, stride 2): GEMM to (dy, dx, cout) columns, then depth-to-space."""
        n, h, w, _ = x.shape
        rows = self.gemm[name].forward(x)
        cp = _pad64(cout)
        fine = (torch.zeros if cp != cout else torch.empty)(n, 2 * h, 2 * w, cp, device=x.device, dtype=x.dtype)
        ops.call("aldi_space_to_depth", rows, fine, n, h, w, 2, cout, cp, rows.shape[3], self.dtc, 1)
        return fine

    def _convt_bwd(self, name, x, cout, dfine, dx=None, accumulate=False):
        n, h, w, _ = x.shape
        lay = self.gemm[name]
        drows = torch.empty(n, h, w, lay.cout_p, device=x.device, dtype=x.dtype)
        ops.call("aldi_space_to_depth", dfine, drows, n, h, w, 2, cout, dfine.shape[3], drows.shape[3], self.dtc, 0)
        return lay.backward(x, drows, dx=dx, accumulate=accumulate)

    def _conv_norm(self, name, x):
        u = self.gemm[name].forward(x)
        y, st = self._ln(u, name + ".norm", self.layout.out_channels)
        return y, (x, u, st)

    def _conv_norm_bwd(self, name, saved, dy, dx=None, accumulate=False):
        x, u, st = saved
        du = self._ln_bwd(u, st, dy, name + ".norm", self.layout.out_channels)
        return self.gemm[name].backward(x, du, dx=dx, accumulate=accumulate)

    # ---- forward (aldi/backbone.py:21-35 + SimpleFeaturePyramid.forward) ----------------------------------------------------
    def forward(self, images_u8, sizes, keep_masks=None, save=True):
        """images_u8 (N, 3, H, W) uint8 on the device, H, W multiples of 32; sizes (N, 2) int32 valid (h, w).
        Returns {"p2", "p3", "p4", "p5"} channels-last; keeps what `backward` needs when save=True."""
        n, _, hp, wp = images_u8.shape
        assert hp % 32 == 0 and wp % 32 == 0
        L, C, dev, dt = self.layout, self.C, images_u8.device, self.dtype
        ps = L.patch
        gh, gw = hp // ps, wp // ps
        keep_masks = keep_masks or [None] * (2 * self.depth)
        S = {"blocks": []} if save else None
        kin = 3 * ps * ps
        patches = torch.empty(n, gh, gw, _pad64(kin), device=dev, dtype=dt)
        ops.call("aldi_patchify_image", images_u8, sizes, patches, n, hp, wp, ps, patches.shape[3], self.dtc,
                 ops.host_floats(self.mean), ops.host_floats(self.std))
        x = self.gemm["patch"].forward(patches)
        ops.call("aldi_add_rows_bcast", x, self._pos_embed(gh, gw), n, gh * gw, C, self.dtc)
        if save:
            S["patches"] = patches
        ws = L.window_size
        for i in range(self.depth):
            p = "net.blocks.%d." % i
            k1, k2 = keep_masks[2 * i], keep_masks[2 * i + 1]
            k1 = k1.to(dev, torch.float32) if k1 is not None else None
            k2 = k2.to(dev, torch.float32) if k2 is not None else None
            ln1, st1 = self._ln(x, p + "norm1", C)
            windowed = i in L.window_blocks
            if windowed:
                nwh, nww = (gh + ws - 1) // ws, (gw + ws - 1) // ws
                win = torch.empty(n * nwh * nww, ws, ws, C, device=dev, dtype=dt)
                ops.call("aldi_window_partition", ln1, win, n, gh, gw, ws, C, self.dtc, 0)
                yw, asave = self._attention(i, win, save)
                y = torch.empty_like(x)
                ops.call("aldi_window_partition", yw, y, n, gh, gw, ws, C, self.dtc, 1)
                del yw, ln1
            else:
                y, asave = self._attention(i, ln1, save)
            x1 = self._residual(y, x, k1)
            ln2, st2 = self._ln(x1, p + "norm2", C)
            hid = self.gemm[p + "fc1"].forward(ln2)
            act = torch.empty_like(hid)
            ops.call("aldi_gelu", hid, None, act, hid.numel(), self.dtc)
            u = self.gemm[p + "fc2"].forward(act)
            x2 = self._residual(u, x1, k2)
            if save:
                S["blocks"].append((x, st1, windowed, asave, k1, x1, st2, ln2, hid, act, k2))
            x = x2
        # ---- SimpleFeaturePyramid: scale factors (4, 2, 1, 0.5) of the stride-16 map -> p2..p5 (p6 is the detector's view)
        outs = {}
        P = {} if save else None
        # p2: ConvT -> LN -> GELU -> ConvT -> 1x1 + LN -> 3x3 + LN
        t0 = self._convt("simfp_2.0", x, C // 2)
        l0, st0 = self._ln(t0, "simfp_2.1", C // 2)
        g0 = torch.empty_like(l0)
        ops.call("aldi_gelu", l0, None, g0, l0.numel(), self.dtc)
        t1 = self._convt("simfp_2.3", g0, C // 4)
        c1, s1 = self._conv_norm("simfp_2.4", t1)
        outs["p2"], s2 = self._conv_norm("simfp_2.5", c1)
        if save:
            P[2] = (t0, st0, l0, g0, s1, s2)
        t0 = self._convt("simfp_3.0", x, C // 2)
        c1, s1 = self._conv_norm("simfp_3.1", t0)
        outs["p3"], s2 = self._conv_norm("simfp_3.2", c1)
        if save:
            P[3] = (s1, s2)
        c1, s1 = self._conv_norm("simfp_4.0", x)
        outs["p4"], s2 = self._conv_norm("simfp_4.1", c1)
        if save:
            P[4] = (s1, s2)
        mp = torch.empty(n, gh // 2, gw // 2, C, device=dev, dtype=dt)
        ops.call("aldi_maxpool2x2", x, mp, self.dtc, n, gh, gw, C)
        c1, s1 = self._conv_norm("simfp_5.1", mp)
        outs["p5"], s2 = self._conv_norm("simfp_5.2", c1)
        if save:
            P[5] = (s1, s2)
            S["pyr"], S["feat"] = P, x
        self.saved = S
        return outs

    # ---- backward ---------------------------------------------------------------------------------------------------
    def backward(self, d_outs):
        """d_outs: {"p2".."p5": gradient of that output (same shape / dtype)}.  Accumulates every parameter gradient into
        `self.grad` (flat, same layout as the parameters)."""
        S = self.saved
        assert S is not None, "forward(save=True) first"
        C, L = self.C, self.layout
        P, feat = S["pyr"], S["feat"]
        n, gh, gw, _ = feat.shape
        dfeat = None

        def into_feat(fn):
            nonlocal dfeat
            if dfeat is None:
                dfeat = torch.empty_like(feat)
                fn(dfeat, False)
            else:
                fn(dfeat, True)

        if "p4" in d_outs:
            s1, s2 = P[4]
            d1 = self._conv_norm_bwd("simfp_4.1", s2, d_outs["p4"])
            into_feat(lambda dx, acc: self._conv_norm_bwd("simfp_4.0", s1, d1, dx=dx, accumulate=acc))
        if "p3" in d_outs:
            s1, s2 = P[3]
            d1 = self._conv_norm_bwd("simfp_3.2", s2, d_outs["p3"])
            dt0 = self._conv_norm_bwd("simfp_3.1", s1, d1)
            into_feat(lambda dx, acc: self._convt_bwd("simfp_3.0", feat, C // 2, dt0, dx=dx, accumulate=acc))
        if "p2" in d_outs:
            t0, st0, l0, g0, s1, s2 = P[2]
            d1 = self._conv_norm_bwd("simfp_2.5", s2, d_outs["p2"])
            dt1 = self._conv_norm_bwd("simfp_2.4", s1, d1)
            dg0 = self._convt_bwd("simfp_2.3", g0, C // 4, dt1)
            dl0 = torch.empty_like(l0)
            ops.call("aldi_gelu", l0, dg0, dl0, l0.numel(), self.dtc)
            dt0 = self._ln_bwd(t0, st0, dl0, "simfp_2.1", C // 2)
            into_feat(lambda dx, acc: self._convt_bwd("simfp_2.0", feat, C // 2, dt0, dx=dx, accumulate=acc))
        if "p5" in d_outs:
            s1, s2 = P[5]
            d1 = self._conv_norm_bwd("simfp_5.2", s2, d_outs["p5"])
            dmp = self._conv_norm_bwd("simfp_5.1", s1, d1)
            dpool = torch.empty_like(feat)
            ops.call("aldi_maxpool2x2_backward", feat, dmp, dpool, self.dtc, n, gh, gw, C)
            if dfeat is None:
                dfeat = dpool
            else:
                self._add(dfeat, dpool)
        assert dfeat is not None, "no gradient reached the backbone"
        dx = dfeat
        ws = L.window_size
        for i in reversed(range(self.depth)):
            p = "net.blocks.%d." % i
            x, st1, windowed, asave, k1, x1, st2, ln2, hid, act, k2 = S["blocks"][i]
            # x2 = x1 + keep2 * fc2(gelu(fc1(norm2(x1))))
            du = self._residual_bwd(dx, k2)
            dact = self.gemm[p + "fc2"].backward(act, du)
            dhid = torch.empty_like(hid)
            ops.call("aldi_gelu", hid, dact, dhid, hid.numel(), self.dtc)
            dln2 = self.gemm[p + "fc1"].backward(ln2, dhid)
            dx1 = self._ln_bwd(x1, st2, dln2, p + "norm2", C, dx=dx)           # accumulates onto the residual gradient
            # x1 = x + keep1 * attn(norm1(x))
            dy = self._residual_bwd(dx1, k1)
            if windowed:
                nwh, nww = (gh + ws - 1) // ws, (gw + ws - 1) // ws
                dyw = torch.empty(n * nwh * nww, ws, ws, C, device=dx.device, dtype=self.dtype)
                ops.call("aldi_window_partition", dy, dyw, n, gh, gw, ws, C, self.dtc, 0)
                dwin = self._attention_bwd(i, asave, dyw)
                dln1 = torch.empty_like(x)
                ops.call("aldi_window_partition", dwin, dln1, n, gh, gw, ws, C, self.dtc, 1)
            else:
                dln1 = self._attention_bwd(i, asave, dy)
            dx = self._ln_bwd(x, st1, dln1, p + "norm1", C, dx=dx1)
            S["blocks"][i] = None
        # x0 = patch_embed(patches) + pos
        ops.call("aldi_sum_over_batch", dx, n, gh * gw, C, self.dtc, self._pos_grad(gh, gw))
        src = self.view("net.pos_embed", self.grad)[C:]
        ops.call("aldi_bicubic_resize", src, L.pos_side, L.pos_side, self._pos_grad(gh, gw), gh, gw, C, C, 1)
        self._pos_grad(gh, gw).zero_()
        self.gemm["patch"].backward(S["patches"], dx, want_dx=False)
        self.saved = None

    def _pos_grad(self, gh, gw):
        key = ("g", gh, gw)
        if key not in self._pos_g:
            self._pos_g[key] = torch.zeros(gh * gw, self.C, device=self.device)
        return self._pos_g[key]

    def _add(self, dst, src):
        """dst += src (activation dtype) through the residual kernel."""
        n, h, w, cp = dst.shape
        ops.call("aldi_layerscale_forward", src, dst, None, None, n * h * w, h * w, self.C, cp, self.dtc, dst)
