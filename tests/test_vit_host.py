"""Host-side logic of the ViTDet path (CPU, no kernels): the flat parameter layouts against the oracle's state dicts
(Detectron2's ViTDet key names and shapes: aldi/backbone.py:37-64, configs/Base-RCNN-VitDetB.yaml), the per-range AdamW
settings against detectron2's rules (aldi/backbone.py:66-84), and the C-ABI struct the attention kernels are called with."""
import ctypes

import torch

from aldi_b200 import arch, vit
from aldi_b200.detector import FlatLayout
from oracle import aldi_ref, vit_ref

SMALL = dict(embed_dim=128, depth=3, num_heads=2, drop_path_rate=0.2, window_block_indexes=(0, 2), lr_decay_rate=0.7)


def test_vit_layout_matches_the_oracle_state_dict_for_b_and_l():
    for size in ("b", "l"):
        lay = vit.ViTLayout(img_size=1024, **vit.vit_config(size))
        ref = vit_ref.build_vitdet_backbone(size).state_dict()
        assert set(lay.entries) == set(ref), (size, set(lay.entries) ^ set(ref))
        for k, (off, n, shape) in lay.entries.items():
            assert tuple(ref[k].shape) == shape and off % 4 == 0, (size, k)
    # ViT-B: 86 M backbone parameters in the ViT proper + the pyramid
    assert 88e6 < vit.ViTLayout(img_size=1024, **vit.vit_config("b")).numel < 95e6


def test_vit_layout_round_trip_keeps_every_tensor_and_orders_the_gemm_operands():
    lay = vit.ViTLayout(img_size=128, **SMALL)
    sd = vit.synthetic_state_dict(lay, seed=1, rel_pos_std=0.1)
    flat = lay.pack(sd)
    back = lay.unpack(flat)
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    # rel_pos_h and rel_pos_w are adjacent: [Rh; Rw] is ONE GEMM operand without a copy
    for i in range(SMALL["depth"]):
        oh, nh, _ = lay.entries["net.blocks.%d.attn.rel_pos_h" % i]
        ow, _, _ = lay.entries["net.blocks.%d.attn.rel_pos_w" % i]
        assert ow == oh + nh
    # ConvTranspose2d weights (cin, cout, 2, 2) are stored as GEMM rows (dy, dx, cout) x cin
    w = sd["simfp_2.0.weight"]
    off, n, _ = lay.entries["simfp_2.0.weight"]
    g = flat[off:off + n].view(2, 2, w.shape[1], w.shape[0])
    assert torch.equal(g[1, 0, 3], w[:, 3, 1, 0])


def test_adamw_ranges_follow_detectron2s_rules():
    """get_adamw_optim(include_vit_lr_decay=True): lr factor 0.7 ** (depth + 1 - layer_id) per parameter
    (get_vit_lr_decay_rate), weight decay 0 on torch.nn.LayerNorm parameters (the blocks' norm1 / norm2 -- NOT detectron2's own
    LayerNorm of the pyramid) and on pos_embed."""
    lay = vit.ViTLayout(img_size=128, **SMALL)
    segs = lay.opt_segments()
    assert segs[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(segs, segs[1:])) and segs[-1][0] + segs[-1][1] == lay.numel
    for k, (off, n, _) in lay.entries.items():
        seg = [s for s in segs if s[0] <= off and off + n <= s[0] + s[1]]
        assert len(seg) == 1, k
        _, _, factor, no_wd = seg[0]
        assert abs(factor - vit_ref.get_vit_lr_decay_rate("backbone." + k, 0.7, SMALL["depth"])) < 1e-12, k
        assert no_wd == (k == "net.pos_embed" or ".norm1." in k or ".norm2." in k), k
    # ViT-L: no layer-wise decay (aldi/trainer.py:206 switches it on for build_vitdet_b_backbone only)
    assert all(s[2] == 1.0 for s in vit.ViTLayout(img_size=1024, **vit.vit_config("l")).opt_segments())


def test_vitdet_detector_layout_matches_the_oracle_detector():
    """Base-RCNN-VitDetB.yaml heads: no FPN parameters, rpn_head.conv.conv{0,1}, box_head.conv{1..4}(+norm), fc1, no fc2."""
    from aldi_b200.train_step import StepConfig, synthetic_state_dict_for
    cfg = StepConfig(backbone="vitdet_b", vit_img_size=128, vit_overrides=dict(SMALL), optimizer="ADAMW")
    sd = synthetic_state_dict_for(cfg, 3)
    net = vit_ref.ViT(img_size=128, embed_dim=128, depth=3, num_heads=2, drop_path_rate=0.2, window_block_indexes=(0, 2))
    model = aldi_ref.ALDI(num_classes=8, backbone=vit_ref.SimpleFeaturePyramid(net), rpn_conv_dims=(-1, -1), box_fc_dims=(1024,),
                          box_conv_dims=(256,) * 4, box_conv_norm="LN")
    model.load_state_dict(sd, strict=True)
    lay = FlatLayout(8, head=arch.VITDET_HEADS)
    keys = {key for (_, _), (_, _, key, _) in lay.entries.items()}
    assert keys == {k for k in sd if not k.startswith("backbone.")}
    assert lay.num_trainable == lay.numel          # no FrozenBN buffers, everything trains
    back = lay.unpack_state_dict(lay.pack_state_dict(sd))
    assert all(torch.equal(back[k], sd[k]) for k in back)
    assert not any(k.startswith("backbone.fpn") or "fc2" in k for k in keys)


def test_attention_params_struct_matches_the_header():
    """ctypes mirror of `aldi_attn_params` (include/aldi_b200.h): field order, and pointer / 64-bit fields on 8-byte offsets."""
    import os
    import re

    from aldi_b200 import lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "aldi_b200.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} aldi_attn_params;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.sub(r"[\s\*]", " ", part).split()[-1])
    assert names == [f[0] for f in lib.AttnParams._fields_], (names, [f[0] for f in lib.AttnParams._fields_])
    for name, ctype in lib.AttnParams._fields_:
        if ctype in (ctypes.c_void_p, ctypes.c_longlong):
            assert getattr(lib.AttnParams, name).offset % 8 == 0, name
