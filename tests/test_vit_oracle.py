"""oracle/vit_ref.py (the ViTDet backbone of aldi/backbone.py:21-64; PARITY UNPINNED -- Detectron2's ViT cannot be
executed here) against independent formulations of each piece: dense relative-position bias through torch's own
scaled_dot_product_attention, per-window attention written out by hand incl. the unmasked zero padding, partition round
trips, Detectron2's module / parameter names, DropPath draw order, the layer-wise lr decay table."""
import math

import torch
import torch.nn.functional as F

from oracle import vit_ref as V


def _dense_bias(q, rel_h, rel_w, H, W):
    """bias[b, (qh, qw), (kh, kw)] = q . rel_h[qh - kh + H - 1] + q . rel_w[qw - kw + W - 1], the slow way."""
    B = q.shape[0]
    qg = q.view(B, H, W, -1)
    bias = torch.zeros(B, H, W, H, W, dtype=q.dtype)
    for qh in range(H):
        for kh in range(H):
            bias[:, qh, :, kh, :] += (qg[:, qh] @ rel_h[qh - kh + H - 1])[:, :, None]
    for qw in range(W):
        for kw in range(W):
            bias[:, :, qw, :, kw] += (qg[:, :, qw] @ rel_w[qw - kw + W - 1])[:, :, None]
    return bias.view(B, H * W, H * W)


def test_attention_matches_sdpa_with_dense_relative_bias():
    torch.manual_seed(0)
    dim, heads, H, W = 32, 4, 5, 5
    att = V.Attention(dim, heads, (H, W)).double()
    with torch.no_grad():
        att.rel_pos_h.normal_()
        att.rel_pos_w.normal_()
    x = torch.randn(2, H, W, dim, dtype=torch.float64)
    got = att(x)
    qkv = att.qkv(x).reshape(2, H * W, 3, heads, -1).permute(2, 0, 3, 1, 4)           # (3, B, heads, HW, d)
    q, k, v = qkv[0], qkv[1], qkv[2]
    bias = _dense_bias(q.reshape(2 * heads, H * W, -1), att.rel_pos_h, att.rel_pos_w, H, W).view(2, heads, H * W, H * W)
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=bias)                    # softmax(q k^T / sqrt(d) + bias) v
    ref = att.proj(ref.permute(0, 2, 1, 3).reshape(2, H, W, dim))
    assert torch.allclose(got, ref, atol=1e-10)


def test_relative_position_table_lookup_and_interpolation():
    t = torch.arange(9, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)      # 2 * 5 - 1 rows
    r = V.get_rel_pos(5, 5, t)
    for qi in range(5):
        for ki in range(5):
            assert float(r[qi, ki, 0]) == qi - ki + 4
    # a table built for 5 positions used on a 9-long axis (non-square input in a global block): linear resampling to 17 rows
    r9 = V.get_rel_pos(9, 9, t)
    want = F.interpolate(t.t()[None], size=17, mode="linear")[0].t()
    assert r9.shape == (9, 9, 3) and torch.equal(r9[0, 8], want[0]) and torch.equal(r9[8, 0], want[16])


def test_window_partition_round_trip_and_padding():
    x = torch.randn(2, 13, 19, 8)
    w, pad_hw = V.window_partition(x, 7)
    assert pad_hw == (14, 21) and w.shape == (2 * 2 * 3, 7, 7, 8)
    assert torch.equal(V.window_unpartition(w, 7, pad_hw, (13, 19)), x)
    assert float(w[1, :, 5:, :].abs().max()) > 0 and float(w[2, :, 5:, :].abs().max()) == 0.0     # right-edge window: zero columns
    assert torch.equal(w[0], x[0, :7, :7])


def test_windowed_block_attends_to_unmasked_zero_padding():
    """Detectron2 pads AFTER norm1 and does not mask: in an edge window the zero tokens are keys (and values) of the
    softmax.  Written out by hand for one edge window."""
    torch.manual_seed(1)
    dim, heads, ws = 16, 2, 4
    owner = type("O", (), {"keep_queue": []})()
    blk = V.Block(dim, heads, 4.0, 0.0, ws, (8, 8), owner).double()
    with torch.no_grad():
        blk.attn.rel_pos_h.normal_()
        blk.attn.rel_pos_w.normal_()
    x = torch.randn(1, 6, 6, dim, dtype=torch.float64)               # 6 x 6 tokens, windows of 4: padded to 8 x 8
    got = blk(x)
    xn = F.pad(blk.norm1(x), (0, 0, 0, 2, 0, 2))                       # zeros AFTER the norm
    win = xn[:, 4:8, 4:8]                                              # bottom-right window: 2 x 2 real, 12 zero tokens
    out = blk.attn(win)[:, :2, :2]
    x1 = x[:, 4:6, 4:6] + out
    ref = x1 + blk.mlp(blk.norm2(x1))
    assert torch.allclose(got[:, 4:6, 4:6], ref, atol=1e-10)
    # and the zero tokens do matter: masking them (-inf on the padded keys) changes the result
    q, k, v = blk.attn.qkv(win).reshape(1, 16, 3, heads, -1).permute(2, 0, 3, 1, 4)
    mask = torch.zeros(4, 4, dtype=torch.bool)
    mask[:2, :2] = True
    bias = _dense_bias(q.reshape(heads, 16, -1), blk.attn.rel_pos_h, blk.attn.rel_pos_w, 4, 4).view(1, heads, 16, 16)
    masked = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.masked_fill(~mask.view(1, 1, 1, 16), -math.inf))
    masked = blk.attn.proj(masked.permute(0, 2, 1, 3).reshape(1, 4, 4, dim))[:, :2, :2]
    assert float((masked - out).detach().abs().max()) > 1e-3


def test_backbone_names_strides_and_gradients():
    torch.manual_seed(2)
    net = V.ViT(img_size=64, patch_size=16, embed_dim=32, depth=3, num_heads=2, drop_path_rate=0.2, window_size=2,
                window_block_indexes=(0, 2), pretrain_img_size=32)
    bb = V.SimpleFeaturePyramid(net, out_channels=16)
    keys = set(bb.state_dict())
    assert {"net.pos_embed", "net.patch_embed.proj.weight", "net.blocks.0.norm1.weight", "net.blocks.1.attn.qkv.bias",
            "net.blocks.1.attn.rel_pos_h", "net.blocks.2.mlp.fc2.weight", "simfp_2.0.weight", "simfp_2.1.weight",
            "simfp_2.3.bias", "simfp_2.4.weight", "simfp_2.4.norm.weight", "simfp_2.5.norm.bias", "simfp_3.0.weight",
            "simfp_3.1.norm.weight", "simfp_4.0.weight", "simfp_4.1.norm.bias", "simfp_5.1.weight", "simfp_5.2.norm.weight"} <= keys
    assert "simfp_2.4.bias" not in keys                                   # conv bias is off when a norm follows
    assert bb.state_dict()["net.pos_embed"].shape == (1, 5, 32)            # 2 x 2 pretraining grid + the class-token slot
    assert bb.state_dict()["net.blocks.0.attn.rel_pos_h"].shape == (3, 16)     # windowed: 2 * 2 - 1
    assert bb.state_dict()["net.blocks.1.attn.rel_pos_w"].shape == (7, 16)     # global: 2 * (64 / 16) - 1
    with torch.no_grad():
        for n, p in bb.named_parameters():
            if "rel_pos" in n or "pos_embed" in n:
                p.normal_(std=0.02)
    x = torch.randn(2, 3, 64, 96)                                         # non-square: 4 x 6 tokens, tables get resampled
    bb.eval()
    out = bb(x)
    assert list(out) == ["p2", "p3", "p4", "p5", "p6"]
    assert [tuple(out[k].shape[2:]) for k in out] == [(16, 24), (8, 12), (4, 6), (2, 3), (1, 2)]
    assert all(out[k].shape[1] == 16 for k in out) and bb._out_feature_strides == {"p2": 4, "p3": 8, "p4": 16, "p5": 32, "p6": 64}
    # training mode: DropPath takes two factors per block with a non-zero rate (block 0 has rate 0), in forward order
    bb.train()
    net.keep_queue = [torch.tensor([1 / 0.9, 0.0]), torch.tensor([0.0, 1 / 0.9]), torch.tensor([1 / 0.8, 1 / 0.8]),
                      torch.tensor([0.0, 0.0])]
    out = bb(x)
    assert net.keep_queue == []
    sum(o.square().mean() for o in out.values()).backward()
    for n, p in bb.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert float(net.blocks[1].attn.qkv.weight.grad.abs().max()) > 0


def test_builders_and_lr_decay_table():
    b = V.build_vitdet_backbone("b")
    assert len(b.net.blocks) == 12 and [i for i, blk in enumerate(b.net.blocks) if blk.window_size == 0] == [2, 5, 8, 11]
    assert b.net.pos_embed.shape == (1, 197, 768) and b.net.blocks[2].attn.rel_pos_h.shape == (127, 64)
    assert abs(b.net.blocks[11].drop_prob - 0.1) < 1e-7 and b.net.blocks[0].drop_prob == 0.0
    # ViT-L geometry only (meta device: no 300 M parameter allocation)
    with torch.device("meta"):
        big = V.build_vitdet_backbone("l")
    assert len(big.net.blocks) == 24 and [i for i, blk in enumerate(big.net.blocks) if blk.window_size == 0] == [5, 11, 17, 23]
    assert big.net.blocks[0].attn.qkv.weight.shape == (3072, 1024)
    f = lambda n: V.get_vit_lr_decay_rate(n, 0.7, 12)                     # noqa: E731  (aldi/backbone.py:74-75)
    assert abs(f("backbone.net.pos_embed") - 0.7 ** 13) < 1e-12 and abs(f("backbone.net.patch_embed.proj.weight") - 0.7 ** 13) < 1e-12
    assert abs(f("backbone.net.blocks.0.attn.qkv.weight") - 0.7 ** 12) < 1e-12
    assert abs(f("backbone.net.blocks.11.mlp.fc2.bias") - 0.7) < 1e-12
    assert f("backbone.simfp_2.0.weight") == 1.0 and f("roi_heads.box_head.fc1.weight") == 1.0


def test_vitdet_detector_heads_and_source_step():
    """Base-RCNN-VitDetB.yaml on the oracle: the simple feature pyramid AS the backbone (no FPN), RPN.CONV_DIMS [-1, -1],
    ROI_BOX_HEAD 4 x conv(256, LN) + one FC -- Detectron2's parameter names, one training-mode step with finite gradients."""
    import random

    from oracle import d2_rcnn as d2
    torch.manual_seed(3)
    random.seed(3)
    net = V.ViT(img_size=64, patch_size=16, embed_dim=32, depth=2, num_heads=2, drop_path_rate=0.0, window_size=2,
                window_block_indexes=(0,), pretrain_img_size=32)
    model = d2.GeneralizedRCNN(num_classes=3, pixel_mean=(123.675, 116.28, 103.53), pixel_std=(58.395, 57.12, 57.375),
                               backbone=V.SimpleFeaturePyramid(net, out_channels=256), rpn_conv_dims=(-1, -1),
                               box_fc_dims=(64,), box_conv_dims=(256, 256, 256, 256), box_conv_norm="LN")
    keys = set(model.state_dict())
    assert {"backbone.net.blocks.1.attn.rel_pos_w", "backbone.simfp_2.4.norm.weight",
            "proposal_generator.rpn_head.conv.conv0.weight", "proposal_generator.rpn_head.conv.conv1.bias",
            "roi_heads.box_head.conv1.weight", "roi_heads.box_head.conv4.norm.bias", "roi_heads.box_head.fc1.weight",
            "roi_heads.box_predictor.cls_score.weight"} <= keys
    assert "roi_heads.box_head.conv1.bias" not in keys and "roi_heads.box_head.fc2.weight" not in keys
    assert model.state_dict()["roi_heads.box_head.fc1.weight"].shape == (64, 256 * 7 * 7)
    assert model.backbone.size_divisibility == 32
    inst = d2.Instances((64, 96))
    inst.gt_boxes = d2.Boxes(torch.tensor([[8.0, 8.0, 40.0, 48.0], [50.0, 10.0, 90.0, 60.0]]))
    inst.gt_classes = torch.tensor([0, 2])
    batch = [{"image": torch.randint(0, 256, (3, 64, 96)).float(), "instances": inst}]
    model.train()
    with d2.EventStorage():
        losses = model(batch)
    assert set(losses) == {"loss_cls", "loss_box_reg", "loss_rpn_cls", "loss_rpn_loc"}
    sum(losses.values()).backward()
    for n, p in model.named_parameters():
        assert p.grad is None or torch.isfinite(p.grad).all(), n
    assert float(model.roi_heads.box_head.conv2.norm.weight.grad.abs().sum()) > 0
    assert float(model.backbone.net.blocks[0].attn.qkv.weight.grad.abs().sum()) > 0


def test_aldi_distillation_step_on_the_vitdet_oracle():
    """The ALDI layer (oracle/aldi_ref.py: pseudo-labeler, hooks, soft losses, aldi/distill.py:115-278) over the ViTDet
    detector variant: eval-mode pseudo-label pass without DropPath, training-mode student / teacher passes with their
    own DropPath draws, the eight distillation loss keys, gradients into the ViT blocks and relative-position tables."""
    import copy
    import random

    from oracle import aldi_ref, d2_rcnn as d2
    torch.manual_seed(4)
    random.seed(4)

    def make():
        net = V.ViT(img_size=64, patch_size=16, embed_dim=32, depth=2, num_heads=2, drop_path_rate=0.2, window_size=2,
                    window_block_indexes=(0,), pretrain_img_size=32)
        m = aldi_ref.ALDI(num_classes=3, pixel_mean=(123.675, 116.28, 103.53), pixel_std=(58.395, 57.12, 57.375),
                          backbone=V.SimpleFeaturePyramid(net, out_channels=256), rpn_conv_dims=(-1, -1), box_fc_dims=(64,),
                          box_conv_dims=(256, 256), box_conv_norm="LN")
        return m, net

    student, net_s = make()
    teacher, net_t = make()
    teacher.load_state_dict(copy.deepcopy(student.state_dict()))
    with torch.no_grad():
        for n, p in list(student.named_parameters()) + list(teacher.named_parameters()):
            if "rel_pos" in n:
                p.normal_(std=0.02)
        teacher.roi_heads.box_predictor.cls_score.bias[0] += 6.0           # a confident teacher: pseudo labels are non-empty
    student.train()
    teacher.train()
    keep = lambda: [torch.tensor([1 / 0.8, 0.0]), torch.tensor([1 / 0.8, 1 / 0.8])]       # noqa: E731  block 1: attn, mlp
    net_s.keep_queue, net_t.keep_queue = keep(), keep()
    dist = aldi_ref.ALDIDistiller(teacher, student, do_hard_cls=False, do_hard_obj=False, do_hard_rpn_reg=False,
                                  do_hard_roi_reg=False, do_cls_dst=True, do_obj_dst=True, do_rpn_reg_dst=True,
                                  do_roih_reg_dst=True, pseudo_label_threshold=0.5)
    img = lambda: {"image": torch.randint(0, 256, (3, 64, 96)).float(), "height": 64, "width": 96}   # noqa: E731
    weak, strong = [img(), img()], [img(), img()]
    with d2.EventStorage():
        losses = aldi_ref.run_model_labeled_unlabeled(student, dist, (None, None, weak, strong), 2, False, lambda l: l.backward())
    assert not net_s.keep_queue and not net_t.keep_queue              # one training-mode pass each; the eval pass draws none
    assert {"loss_cls_distill", "loss_box_reg_distill", "loss_rpn_cls_distill", "loss_rpn_loc_distill", "loss_obj_bce_distill",
            "loss_rpn_l1_distill", "loss_cls_ce_distill", "loss_roih_l1_distill"} <= set(losses)
    assert all(torch.isfinite(torch.as_tensor(v)) for v in losses.values())
    assert sum(len(d["instances"]) for d in weak) > 0                  # pseudo labels were attached to the weak dicts
    assert float(net_s.blocks[1].attn.rel_pos_h.grad.abs().sum()) > 0 and float(net_s.patch_embed.proj.weight.grad.abs().sum()) > 0
    assert all(p.grad is None for p in teacher.parameters())
