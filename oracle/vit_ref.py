"""TEST INFRASTRUCTURE ONLY — CPU restatement of the ViTDet backbone the reference builds in aldi/backbone.py:21-64.

Only tests/ may import this.  **Parity: the ViT trunk is PINNED to an independent published port, the pyramid is UNPINNED.**
`build_vitdet_b_backbone` / `build_vitdet_l_backbone` instantiate Detectron2's `SimpleFeaturePyramid(net=ViT(...))` from the
LazyConfig `common/models/mask_rcnn_vitdet.py`; that code lives in the third-party `detectron2` package (`pyproject.toml:18`, a
git dependency with no pinned commit) which is absent from /root/reference and cannot be installed here, and the reference holds
no golden vectors for it.  What is restated below is detectron2 v0.6's `modeling/backbone/vit.py` + `modeling/backbone/utils.py`
as published, with the reference's own changes applied on top:
  * `square_pad = 0` (aldi/backbone.py:40,48) -- images are padded to the size divisibility only, not to a square;
  * `checkpointed_vit_forward` (aldi/backbone.py:21-35): patch_embed -> + get_abs_pos(pos_embed) -> blocks -> NCHW,
    with per-block activation checkpointing (`VIT.USE_ACT_CHECKPOINT`), which changes memory, not values;
  * ViT-L: embed_dim 1024, depth 24, 16 heads, drop_path 0.4, global attention in blocks 5, 11, 17, 23 (:50-58).
The pin: `ViT` (patch embedding, bicubic pos_embed, windowed / global blocks, decomposed relative position with resampled
tables, zero-padded windows) reproduces HuggingFace transformers' `VitDetModel` -- a port of the same detectron2 file made by
other hands -- to 1e-12 in float64 on outputs and every parameter gradient (tests/golden/make_vit_golden.py executes it,
tests/test_vit_golden.py replays the committed vectors and, when transformers is importable, the live model).
`SimpleFeaturePyramid`, `get_vit_lr_decay_rate` and the ViTDet heads in d2_rcnn.py have no such counterpart: restated only;
tests/test_vit_oracle.py checks them against independent formulations (dense relative-position bias, torch's
scaled_dot_product_attention, partition round trips).

Module tree and parameter names are Detectron2's (`net.pos_embed`, `net.patch_embed.proj`, `net.blocks.{i}.attn.qkv`,
`...attn.rel_pos_h`, `simfp_{2..5}.{k}`, `...norm.weight`), so released ViTDet checkpoints would load with strict=True.
Parity traps reproduced: (1) window padding is added AFTER norm1 and is NOT masked -- the zero tokens take part in the
softmax of their window as keys; (2) the decomposed relative-position term uses the UNSCALED q; (3) relative-position
tables whose length differs from 2 * max(q, k) - 1 (non-square inputs in global blocks) are linearly interpolated;
(4) `pos_embed` holds a class-token slot that is dropped, and is bicubically resampled when the token grid is not 14 x 14
(`pretrain_img_size` 224).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class LayerNormCF(nn.Module):
    """detectron2.layers.batch_norm.LayerNorm: over the channel axis of NCHW, eps 1e-6 (`get_norm("LN", c)`)."""

    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight, self.bias = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c))
        self.eps = eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class ConvNorm(nn.Conv2d):
    """detectron2.layers.Conv2d: conv -> norm (-> activation, none here); the norm is the submodule `.norm`."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__(cin, cout, k, padding=padding, bias=False)     # bias = (norm == "") = False for "LN"
        self.norm = LayerNormCF(cout)

    def forward(self, x):
        return self.norm(super().forward(x))


def window_partition(x, ws):
    """(B, H, W, C) -> (B * nW, ws, ws, C), zero padding at the bottom / right to a multiple of ws."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(windows, ws, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = windows.shape[0] // (Hp * Wp // ws // ws)
    x = windows.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    return x[:, :H, :W, :].contiguous() if (Hp > H or Wp > W) else x


def get_rel_pos(q_size, k_size, rel_pos):
    """rows of the (2 * max - 1, head_dim) table for every (query, key) coordinate pair: (q_size, k_size, head_dim)."""
    max_rel = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel:
        r = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel, mode="linear")
        r = r.reshape(-1, max_rel).permute(1, 0)
    else:
        r = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return r[rel.long()]


def add_decomposed_rel_pos(attn, q, rel_pos_h, rel_pos_w, q_size, k_size):
    """attn (B, qh*qw, kh*kw) += q . Rh[qh - kh] + q . Rw[qw - kw]   (MViTv2's decomposition; q is NOT scaled)."""
    qh, qw = q_size
    kh, kw = k_size
    Rh, Rw = get_rel_pos(qh, kh, rel_pos_h), get_rel_pos(qw, kw, rel_pos_w)
    B, _, dim = q.shape
    rq = q.reshape(B, qh, qw, dim)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    return (attn.view(B, qh, qw, kh, kw) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(B, qh * qw, kh * kw)


class Attention(nn.Module):
    def __init__(self, dim, heads, input_size, use_rel_pos=True):
        super().__init__()
        self.heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv, self.proj = nn.Linear(dim, dim * 3, bias=True), nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:                                                  # rel_pos_zero_init=True
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, dim // heads))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, dim // heads))

    def forward(self, x):
        B, H, W, _ = x.shape
        qkv = self.qkv(x).reshape(B, H * W, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.reshape(3, B * self.heads, H * W, -1).unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        if self.use_rel_pos:
            attn = add_decomposed_rel_pos(attn, q, self.rel_pos_h, self.rel_pos_w, (H, W), (H, W))
        attn = attn.softmax(dim=-1)
        x = (attn @ v).view(B, self.heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
        return self.proj(x)


class Mlp(nn.Module):            # timm.layers.Mlp, drop = 0
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, drop_path, window_size, input_size, owner):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads, input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.window_size, self.drop_prob = window_size, drop_path
        self._owner = [owner]

    def drop_path(self, x):
        """timm DropPath: per-sample keep / keep_prob in training; the factors come from the owner's `keep_queue`
        (filled by the test) instead of `bernoulli_`, two draws per block in forward order."""
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = self._owner[0].keep_queue.pop(0).to(x.dtype)
        return x * keep.view(-1, *([1] * (x.dim() - 1)))

    def forward(self, x):
        shortcut = x
        x = self.norm1(x)
        if self.window_size > 0:
            H, W = x.shape[1], x.shape[2]
            x, pad_hw = window_partition(x, self.window_size)
        x = self.attn(x)
        if self.window_size > 0:
            x = window_unpartition(x, self.window_size, pad_hw, (H, W))
        x = shortcut + self.drop_path(x)
        return x + self.drop_path(self.mlp(self.norm2(x)))


class PatchEmbed(nn.Module):
    def __init__(self, patch, cin, dim):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)


def get_abs_pos(abs_pos, has_cls_token, hw):
    h, w = hw
    if has_cls_token:
        abs_pos = abs_pos[:, 1:]
    size = int(math.sqrt(abs_pos.shape[1]))
    assert size * size == abs_pos.shape[1]
    if size != h or size != w:
        p = F.interpolate(abs_pos.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(h, w), mode="bicubic",
                          align_corners=False)
        return p.permute(0, 2, 3, 1)
    return abs_pos.reshape(1, h, w, -1)


class ViT(nn.Module):
    def __init__(self, img_size=1024, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 drop_path_rate=0.1, window_size=14, window_block_indexes=(0, 1, 3, 4, 6, 7, 9, 10),
                 pretrain_img_size=224):
        super().__init__()
        self.patch_embed = PatchEmbed(patch_size, 3, embed_dim)
        n_pos = (pretrain_img_size // patch_size) ** 2 + 1               # pretrain_use_cls_token=True
        self.pos_embed = nn.Parameter(torch.zeros(1, n_pos, embed_dim))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth, device="cpu")]
        grid = (img_size // patch_size, img_size // patch_size)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, dpr[i],
                                           window_size if i in window_block_indexes else 0, grid, self)
                                     for i in range(depth)])
        self.keep_queue = []
        self.embed_dim, self.patch_size = embed_dim, patch_size

    def forward(self, x):
        """aldi/backbone.py:21-35 (checkpointing elided: it does not change values)."""
        x = self.patch_embed(x)
        x = x + get_abs_pos(self.pos_embed, True, (x.shape[1], x.shape[2]))
        for blk in self.blocks:
            x = blk(x)
        return {"last_feat": x.permute(0, 3, 1, 2)}


class SimpleFeaturePyramid(nn.Module):
    """ViTDet's pyramid from the single stride-16 map: scale factors (4, 2, 1, 0.5) -> p2..p5, p6 = LastLevelMaxPool."""

    def __init__(self, net, out_channels=256, scale_factors=(4.0, 2.0, 1.0, 0.5)):
        super().__init__()
        self.net = net
        dim = net.embed_dim
        self.stage_names = []
        for scale in scale_factors:
            out_dim = dim
            if scale == 4.0:
                layers = [nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2), LayerNormCF(dim // 2), nn.GELU(),
                          nn.ConvTranspose2d(dim // 2, dim // 4, kernel_size=2, stride=2)]
                out_dim = dim // 4
            elif scale == 2.0:
                layers = [nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2)]
                out_dim = dim // 2
            elif scale == 1.0:
                layers = []
            elif scale == 0.5:
                layers = [nn.MaxPool2d(kernel_size=2, stride=2)]
            else:
                raise NotImplementedError(scale)
            layers += [ConvNorm(out_dim, out_channels, 1), ConvNorm(out_channels, out_channels, 3, padding=1)]
            stage = int(math.log2(net.patch_size / scale))
            self.add_module("simfp_%d" % stage, nn.Sequential(*layers))
            self.stage_names.append("simfp_%d" % stage)
        self._out_features = ["p%d" % int(math.log2(net.patch_size / s)) for s in scale_factors] + ["p6"]
        self._out_feature_strides = {f: 2 ** int(f[1:]) for f in self._out_features}
        self._out_feature_channels = {f: out_channels for f in self._out_features}
        self.size_divisibility = 32                                      # strides[-1]; square_pad = 0 (aldi/backbone.py:40)

    def forward(self, x):
        feat = self.net(x)["last_feat"]
        res = [getattr(self, n)(feat) for n in self.stage_names]
        res.append(F.max_pool2d(res[-1], kernel_size=1, stride=2, padding=0))          # LastLevelMaxPool on p5
        return dict(zip(self._out_features, res))


def build_vitdet_backbone(size="b"):
    """aldi/backbone.py:37-64."""
    if size == "b":
        net = ViT()
    elif size == "l":
        net = ViT(embed_dim=1024, depth=24, num_heads=16, drop_path_rate=0.4,
                  window_block_indexes=tuple(list(range(0, 5)) + list(range(6, 11)) + list(range(12, 17)) + list(range(18, 23))))
    else:
        raise ValueError(size)
    return SimpleFeaturePyramid(net)


def get_vit_lr_decay_rate(name, lr_decay_rate=1.0, num_layers=12):
    """detectron2.modeling.backbone.vit.get_vit_lr_decay_rate (layer-wise lr decay; on for ViT-B only:
    aldi/trainer.py:206, aldi/backbone.py:74-79 -- 0.7 over 12 layers)."""
    layer_id = num_layers + 1
    if name.startswith("backbone"):
        if ".pos_embed" in name or ".patch_embed" in name:
            layer_id = 0
        elif ".blocks." in name and ".residual." not in name:
            layer_id = int(name[name.find(".blocks."):].split(".")[2]) + 1
    return lr_decay_rate ** (num_layers + 1 - layer_id)
