"""Selection kernels against the oracle's Detectron2 primitives on IDENTICAL inputs: index sets must be exact."""
import ctypes

import pytest
import torch

import parity_utils as pu
from oracle import d2_rcnn as d2

pytestmark = pytest.mark.gpu


def _levels(shapes):
    from aldi_b200 import ops
    from aldi_b200.detector import FPN_STRIDES, SCALE_CLAMP, cell_anchors
    return ops.make_rpn_levels(shapes, FPN_STRIDES, cell_anchors(), 16, SCALE_CLAMP, 0.0)


def _anchors(shapes):
    gen = d2.DefaultAnchorGenerator()
    return gen([torch.zeros(1, 1, h, w) for h, w in shapes])


def _rand_gt(gen, n, h, w, count):
    out = []
    for _ in range(n):
        k = count
        xy = torch.rand(k, 2, generator=gen) * torch.tensor([w * 0.7, h * 0.7])
        wh = 8 + torch.rand(k, 2, generator=gen) * torch.tensor([w * 0.3, h * 0.3])
        out.append(torch.cat([xy, xy + wh], 1))
    return out


@pytest.mark.parametrize("count", [0, 1, 7, 40])
def test_rpn_label_and_sample_exact(count):
    from aldi_b200 import ops, sampling
    from aldi_b200 import lib as _l
    shapes = [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)]
    H, W, n = 128, 160, 3
    gen = torch.Generator().manual_seed(count)
    gts = _rand_gt(gen, n, H, W, count)
    if count == 7:
        gts[1][0] = torch.tensor([0., 0., 0., 0.])       # degenerate GT: IoU 0 with every anchor (D2 low-quality quirk)
    lv = _levels(shapes)
    seed, pass_id = 12345, 3
    pu.install_device_sampler({pass_id: seed})
    d2.SAMPLE_CTX["pass"] = pass_id
    rpn = d2.RPN()
    insts = [d2.Instances((H, W), gt_boxes=d2.Boxes(g), gt_classes=torch.zeros(len(g), dtype=torch.int64)) for g in gts]
    ref_labels, ref_boxes = rpn.label_and_sample_anchors(_anchors(shapes), insts)
    d2.set_sample_chooser(None)
    gmax = 64
    gb = torch.zeros(n, gmax, 4)
    cnt = torch.zeros(n, dtype=torch.int32)
    for i, g in enumerate(gts):
        gb[i, :len(g)] = g
        cnt[i] = len(g)
    dev = "cuda"
    total = lv.total_locs * 3
    labels = torch.empty(n, total, dtype=torch.int8, device=dev)
    matched = torch.empty(n, total, dtype=torch.int32, device=dev)
    stats = torch.zeros(n, 2, dtype=torch.int32, device=dev)
    wsb = int(_l.load().aldi_rpn_label_workspace_bytes(n, gmax))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    salts = torch.tensor([sampling.make_salt(pass_id, sampling.SITE_RPN, i) for i in range(n)], dtype=torch.int32, device=dev)
    seed_t = torch.tensor([seed], dtype=torch.int32, device=dev)   # device-resident seed (graph-replayable)
    ops.call("aldi_rpn_label_anchors", ctypes.byref(lv), n, gb.to(dev), cnt.to(dev), gmax, 0.3, 0.7, 256, 0.5, seed_t, salts,
             ws, wsb, labels, matched, stats)
    torch.cuda.synchronize()
    for i in range(n):
        got, want = labels[i].cpu(), ref_labels[i]
        diff = (got != want).nonzero().flatten()
        assert diff.numel() == 0, ("image", i, "mismatches", diff.numel(), diff[:10].tolist(), got[diff[:10]].tolist(),
                                   want[diff[:10]].tolist(), stats.cpu().tolist())
        if count:
            pos = (want == 1).nonzero().flatten()
            mb = gb[i][matched[i].cpu().long()[pos]]
            assert torch.equal(mb, ref_boxes[i][pos]), "matched gt boxes of positive anchors"


def test_roi_label_sample_exact():
    from aldi_b200 import ops, sampling
    H, W, n, P = 128, 160, 2, 300
    gen = torch.Generator().manual_seed(5)
    gts = _rand_gt(gen, n, H, W, 6)
    props = _rand_gt(gen, n, H, W, P)
    for i in range(n):  # make some proposals overlap the gt well
        props[i][:6] = gts[i] + torch.randn(6, 4, generator=gen) * 2
    classes = [torch.randint(0, 8, (6,), generator=gen) for _ in range(n)]
    seed, pass_id = 777, 100
    pu.install_device_sampler({pass_id: seed})
    d2.SAMPLE_CTX["pass"] = pass_id
    heads = d2.StandardROIHeads(num_classes=8)
    plist = [d2.Instances((H, W), proposal_boxes=d2.Boxes(p), objectness_logits=torch.zeros(P)) for p in props]
    tlist = [d2.Instances((H, W), gt_boxes=d2.Boxes(g), gt_classes=c) for g, c in zip(gts, classes)]
    with d2.EventStorage():
        ref = heads.label_and_sample_proposals(plist, tlist)
    d2.set_sample_chooser(None)
    dev = "cuda"
    gmax, S = 32, 512
    gb = torch.zeros(n, gmax, 4); gc = torch.zeros(n, gmax, dtype=torch.int32); cnt = torch.full((n,), 6, dtype=torch.int32)
    for i in range(n):
        gb[i, :6] = gts[i]; gc[i, :6] = classes[i].int()
    pb = torch.stack(props).to(dev)
    pc = torch.full((n,), P, dtype=torch.int32, device=dev)
    m = n * S
    rois = torch.empty(m, 4, device=dev); rgt = torch.empty(m, 4, device=dev)
    rb = torch.empty(m, dtype=torch.int32, device=dev); rc = torch.empty(m, dtype=torch.int32, device=dev)
    rs = torch.empty(m, dtype=torch.int32, device=dev); rcount = torch.zeros(n, dtype=torch.int32, device=dev)
    salts = torch.tensor([sampling.make_salt(pass_id, sampling.SITE_ROI, i) for i in range(n)], dtype=torch.int32, device=dev)
    seed_t = torch.tensor([seed], dtype=torch.int32, device=dev)
    ops.call("aldi_roi_label_sample", pb, pc, P, n, gb.to(dev), gc.to(dev), cnt.to(dev), gmax, 0.5, 8, S, 0.25, seed_t, salts, 1,
             rois, rb, rc, rgt, rs, rcount, None)
    torch.cuda.synchronize()
    for i in range(n):
        c = int(rcount[i])
        assert c == len(ref[i]), (c, len(ref[i]))
        sl = slice(i * S, i * S + c)
        assert torch.equal(rois[sl].cpu(), ref[i].proposal_boxes.tensor), "sampled boxes (order: fg asc, bg asc)"
        assert torch.equal(rc[sl].cpu().long(), ref[i].gt_classes), "assigned classes"
        fg = ref[i].gt_classes < 8
        assert torch.equal(rgt[sl].cpu()[fg], ref[i].gt_boxes.tensor[fg]), "matched gt boxes"
        assert bool((rc[i * S + c:(i + 1) * S] == -1).all()), "padding rows"


@pytest.mark.parametrize("training", [True, False])
def test_rpn_proposals_exact(training):
    from aldi_b200.detector import Detector
    shapes = [(64, 80), (32, 40), (16, 20), (8, 10), (4, 5)]
    n, H, W = 2, 256, 320
    gen = torch.Generator().manual_seed(3)
    det = Detector(8)
    lv = _levels(shapes)
    rpn_out = torch.zeros(n, lv.total_locs, 16)
    rpn_out[..., :3] = torch.randn(n, lv.total_locs, 3, generator=gen) * 2
    rpn_out[..., 3:15] = torch.randn(n, lv.total_locs, 12, generator=gen) * 0.5
    logits, deltas = pu.rpn_out_to_d2(rpn_out, lv, n)
    rpn = d2.RPN()
    rpn.train(training)
    anchors = _anchors(shapes)
    lg = [s.permute(0, 2, 3, 1).flatten(1) for s in logits]
    dl = [x.view(n, -1, 4, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2) for x in deltas]
    ref = rpn.predict_proposals(anchors, lg, dl, [(H, W - 7), (H - 30, W)])
    sizes = torch.tensor([[H, W - 7], [H - 30, W]], dtype=torch.int32, device="cuda")
    pre = 2000 if training else 1000
    out = det.proposals(rpn_out.cuda(), lv, sizes, pre, 1000, 0.7)
    torch.cuda.synchronize()
    for i in range(n):
        c = int(out["count"][i])
        got_s, want_s = out["scores"][i, :c].cpu(), ref[i].objectness_logits
        m = min(c, len(want_s))
        neq = (got_s[:m] != want_s[:m]).nonzero().flatten()
        first = int(neq[0]) if neq.numel() else -1
        info = (i, "count", c, len(want_s), "first mismatch", first,
                got_s[max(first - 2, 0):first + 3].tolist(), want_s[max(first - 2, 0):first + 3].tolist(),
                "n mismatches", neq.numel())
        assert c == len(ref[i]) and neq.numel() == 0, info
        gb, wb = out["boxes"][i, :c].cpu(), ref[i].proposal_boxes.tensor
        # Proposals with EQUAL objectness: the device orders them by anchor index; the reference order is whatever
        # torch.topk / sort return for ties on the model device (differs between CPU and CUDA builds of the reference
        # itself).  Compare each group of equal scores as a set; everything else position by position.
        gb, wb = _canonical_tie_order(got_s, gb), _canonical_tie_order(want_s, wb)
        bad = ((gb - wb).abs() > 1e-3 + 1e-5 * wb.abs()).any(1).nonzero().flatten()
        assert bad.numel() == 0, ("boxes", i, bad.numel(), bad[:4].tolist(), gb[bad[:4]].tolist(), wb[bad[:4]].tolist(),
                                  got_s[bad[:4]].tolist(), out["cats"][i, bad[:4]].tolist())


def _canonical_tie_order(scores, boxes):
    """Within runs of equal scores (the list is score-descending) order the boxes lexicographically."""
    boxes = boxes.clone()
    n, a = len(scores), 0
    while a < n:
        b = a + 1
        while b < n and scores[b] == scores[a]:
            b += 1
        if b - a > 1:
            rows = sorted(boxes[a:b].tolist())
            boxes[a:b] = torch.tensor(rows)
        a = b
    return boxes


def test_roi_inference_exact():
    from aldi_b200.detector import Detector
    n, P, K, H, W = 2, 200, 8, 256, 320
    gen = torch.Generator().manual_seed(9)
    det = Detector(K)
    props = torch.stack(_rand_gt(gen, n, H, W, P))
    pred = torch.zeros(n * P, 64)
    pred[:, :K + 1] = torch.randn(n * P, K + 1, generator=gen) * 3
    pred[:, K + 1:5 * K + 1] = torch.randn(n * P, 4 * K, generator=gen)
    layer = d2.FastRCNNOutputLayers(num_classes=K)
    plist = [d2.Instances((H, W), proposal_boxes=d2.Boxes(props[i])) for i in range(n)]
    ref, _ = layer.inference((pred[:, :K + 1], pred[:, K + 1:5 * K + 1]), plist)
    pd = {"boxes": props.cuda(), "scores": torch.zeros(n, P, device="cuda"),
          "count": torch.full((n,), P, dtype=torch.int32, device="cuda")}
    sizes = torch.tensor([[H, W]] * n, dtype=torch.int32, device="cuda")
    for thr in (0.05, 0.8):
        out = det.detections(pred.cuda(), pd, sizes, thr)
        torch.cuda.synchronize()
        for i in range(n):
            keep = ref[i].scores > thr
            c = int(out["count"][i])
            want = ref[i].scores[keep]
            if thr == 0.05:
                assert c == len(ref[i])
            # with a higher candidate threshold the surviving detections above it are unchanged (greedy NMS)
            assert torch.allclose(out["scores"][i, :len(want)].cpu(), want, rtol=1e-5), (thr, c, len(want))
            assert torch.equal(out["cats"][i, :len(want)].cpu().long(), ref[i].pred_classes[keep])
            assert torch.allclose(out["boxes"][i, :len(want)].cpu(), ref[i].pred_boxes.tensor[keep], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("seed,quant", [(0, 0.0), (1, 0.25)])
def test_segmented_nms_equals_sorted_nms(seed, quant):
    """aldi_nms_segmented (per-level compaction / mask / scan in parallel + rank merge) must reproduce aldi_nms_sorted
    bit for bit on level-major, per-level score-sorted candidates — including invalid candidates, score ties across and
    within levels (quantised scores) and the keep[:post_topk] cut."""
    from aldi_b200.detector import Detector
    g = torch.Generator().manual_seed(seed)
    n, lens = 3, [2000, 2000, 1200, 300, 45]
    offs = [sum(lens[:i]) for i in range(len(lens))]
    stride = sum(lens)
    cb = torch.zeros(n, stride, 4)
    cs = torch.zeros(n, stride)
    cc = torch.zeros(n, stride, dtype=torch.int32)
    for i in range(n):
        for l, (o, k) in enumerate(zip(offs, lens)):
            s = torch.randn(k, generator=g) * 2
            if quant:
                s = (s / quant).round() * quant
            s = s.sort(descending=True).values
            ctr = torch.rand(k, 2, generator=g) * 600
            wh = torch.rand(k, 2, generator=g) * 120 * (l + 1) + 4
            cb[i, o:o + k] = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
            cs[i, o:o + k] = s
            cc[i, o:o + k] = l
    cv = (torch.rand(n, stride, generator=g) > 0.03).to(torch.uint8)
    cb, cs, cc, cv = cb.cuda(), cs.cuda(), cc.cuda(), cv.cuda()
    for post in (1000, 137):
        a = Detector.nms(cb, cs, cc, cv, None, 0.7, post)
        b = Detector.nms_segmented(cb, cs, cv, offs, lens, 0.7, post)
        torch.cuda.synchronize()
        assert torch.equal(a["count"], b["count"])
        for i in range(n):
            c = int(a["count"][i])
            assert c > 0
            for k in ("src", "cats", "scores", "boxes"):
                assert torch.equal(a[k][i, :c], b[k][i, :c]), (post, i, k)


def test_conv_halo_path_matches_fp32_kernel():
    """conv_tc's halo-tile mode (3x3, 64 -> 64 channels: nine shifted UMMA operands inside one shared-memory tile,
    resident weights) against the fp32 CUDA-core kernel, on sizes that are not multiples of the 16 x 8 tile."""
    from aldi_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for n, h, w in ((2, 48, 72), (1, 37, 50), (3, 16, 8)):
        x = torch.randn(n, h, w, 64, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(64, 9 * 64, device="cuda", generator=g) / 24.0).bfloat16()
        sc = torch.rand(64, device="cuda", generator=g) + 0.5
        bi = torch.randn(64, device="cuda", generator=g)
        out = torch.zeros(n, h, w, 64, device="cuda", dtype=torch.bfloat16)
        ref = torch.zeros(n, h, w, 64, device="cuda", dtype=torch.float32)
        ops.conv(x, wt, out, taps_h=3, taps_w=3, pad_h=1, pad_w=1, scale=sc, bias=bi, relu=True)
        ops.conv(x.float(), wt.float(), ref, taps_h=3, taps_w=3, pad_h=1, pad_w=1, scale=sc, bias=bi, relu=True)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max() / ref.abs().max()
        assert float(err) < 1e-2, (n, h, w, float(err))


@pytest.mark.parametrize("cin,cout,k", [(128, 128, 3), (128, 512, 1), (256, 256, 3), (128, 128, 1)])
def test_wgrad_tc_matches_fp32_kernel(cin, cout, k):
    """wgrad_tc (incl. the tap-paired N = 256 mode for 128 input channels) against the fp32 CUDA-core kernel."""
    from aldi_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    n, h, w = 2, 40, 56
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    dy = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
    dw = torch.zeros(cout, k * k * cin, device="cuda")
    ref = torch.zeros(cout, k * k * cin, device="cuda")
    ops.wgrad(x, dy, dw, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2)
    ops.wgrad(x.float(), dy.float(), ref, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2)
    torch.cuda.synchronize()
    err = (dw - ref).abs().max() / ref.abs().max()
    assert float(err) < 5e-3, float(err)


@pytest.mark.parametrize("cin,cout,k,hw", [(256, 256, 3, (40, 56)), (1024, 256, 1, (24, 40)), (256, 64, 1, (64, 64)),
                                           (128, 128, 3, (33, 47)), (12544, 1024, 1, (1, 300))])
def test_wgrad_fused_bias_gradient(cin, cout, k, hw):
    """aldi_wgrad_tc(dbias=): the bias gradient summed from the dy stages inside the weight-gradient kernel must equal
    aldi_colsum (incl. split-K items, several cout tiles, ragged pixel tiles, the box head's (1,1,M,C) layout)."""
    from aldi_b200 import lib as _l, ops
    g = torch.Generator(device="cuda").manual_seed(2)
    n, (h, w) = 2, hw
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    dy = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
    store = cout if cout != 64 else 41                        # predictor-like: 41 of 64 padded channels
    dw = torch.zeros(store, k * k * cin, device="cuda")
    db = torch.zeros(store, device="cuda")
    ops.wgrad(x, dy, dw, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2, cout_store=store, dbias=db)
    ref = torch.zeros(store, device="cuda")
    ops.call("aldi_colsum", dy, _l.BF16, 1, n * h * w, 0, cout, store, 1.0, ref)
    torch.cuda.synchronize()
    assert torch.allclose(db, ref, rtol=1e-4, atol=1e-2), float((db - ref).abs().max())
    assert float(ref.abs().max()) > 1.0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_roi_align_forward_backward_vs_torchvision(seed):
    """Multi-level RoIAlign (forward and the pixel-centric backward) against detectron2's ROIPooler restatement on
    torchvision.ops.roi_align: random boxes of every size, some partly or wholly outside the image, tiny boxes."""
    from aldi_b200 import ops
    g = torch.Generator().manual_seed(seed)
    n, c = 2, 64
    shapes = [(40, 48), (20, 24), (10, 12), (5, 6)]
    feats = [torch.randn(n, c, h, w, generator=g) for h, w in shapes]
    m = 96
    ctr = torch.rand(m, 2, generator=g) * torch.tensor([192.0, 160.0])
    wh = torch.exp(torch.rand(m, 2, generator=g) * 5.5) * 1.5          # 1.5 .. 370 px
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    boxes[:8] += 120.0                                                  # some far outside
    boxes[8:12, 2:] = boxes[8:12, :2] + 0.3                             # tiny
    batch = torch.randint(0, n, (m,), generator=g)
    order = torch.argsort(batch, stable=True)
    boxes, batch = boxes[order], batch[order]
    dout = torch.randn(m, c, 7, 7, generator=g)
    fr = [f.clone().requires_grad_(True) for f in feats]
    pooler = d2.ROIPooler()
    out_ref = pooler(fr, [d2.Boxes(boxes[batch == i]) for i in range(n)])
    out_ref.backward(dout)
    fd = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats]
    out = torch.empty(m, 7, 7, c, device="cuda")
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ops.roi_align(fd, boxes.cuda(), batch.int().cuda(), out=out, scales=scales)
    dfe = [torch.zeros_like(f) for f in fd]
    ops.roi_align(fd, boxes.cuda(), batch.int().cuda(), dout=dout.permute(0, 2, 3, 1).contiguous().cuda(), dfeats=dfe, scales=scales)
    torch.cuda.synchronize()
    assert torch.allclose(out.cpu().permute(0, 3, 1, 2), out_ref.detach(), rtol=1e-4, atol=1e-5)
    for l in range(4):
        got, want = dfe[l].cpu().permute(0, 3, 1, 2), fr[l].grad
        err = (got - want).abs().max() / (want.abs().max() + 1e-12)
        assert float(err) < 1e-4, (l, float(err))
