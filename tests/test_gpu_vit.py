"""ViTDet backbone on the device vs oracle/vit_ref.py (BASELINE configs[2]; the oracle restates detectron2's
vit.py + aldi/backbone.py:21-64 and is parity-unpinned, see its header): the small kernels, the attention op in both
arithmetic modes (CUDA-core fp32 = parity mode, tcgen05 bf16 = the benchmarked mode) and the whole backbone forward +
backward (pyramid outputs and every parameter gradient)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


# ---------------------------------------------------------------------------------------------------------------------
# small kernels
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_window_partition_round_trip_matches_oracle(dtype):
    from aldi_b200 import lib as _l, ops
    from oracle import vit_ref
    dev = _dev()
    n, h, w, c, ws = 2, 16, 20, 64, 14
    x = torch.randn(n, h, w, c).to(dtype)
    ref, (hp, wp) = vit_ref.window_partition(x.float(), ws)
    dtc = _l.BF16 if dtype == torch.bfloat16 else _l.F32
    xd = x.to(dev)
    win = torch.full((ref.shape[0], ws, ws, c), 7.0, device=dev, dtype=dtype)
    ops.call("aldi_window_partition", xd, win, n, h, w, ws, c, dtc, 0)
    assert torch.equal(win.float().cpu(), ref.to(dtype).float())
    back = torch.zeros_like(xd)
    ops.call("aldi_window_partition", win, back, n, h, w, ws, c, dtc, 1)
    assert torch.equal(back.cpu(), x)


def test_bicubic_and_linear_resampling_match_torch():
    from aldi_b200 import ops
    dev = _dev()
    c = 24
    src = torch.randn(14, 14, c)
    for dh, dw in ((16, 20), (14, 14), (5, 9), (64, 64)):
        ref = F.interpolate(src.permute(2, 0, 1)[None], size=(dh, dw), mode="bicubic", align_corners=False)[0].permute(1, 2, 0)
        out = torch.empty(dh * dw, c, device=dev)
        ops.call("aldi_bicubic_resize", src.to(dev), 14, 14, out, dh, dw, c, c, 0)
        assert _rel(out.cpu().view(dh, dw, c), ref) < 1e-5, (dh, dw)
        # transpose: <A x, g> == <x, A^T g>
        g = torch.randn(dh * dw, c)
        gsrc = torch.zeros(14, 14, c, device=dev)
        ops.call("aldi_bicubic_resize", gsrc, 14, 14, g.to(dev), dh, dw, c, c, 1)
        s = src.clone().requires_grad_(True)
        F.interpolate(s.permute(2, 0, 1)[None], size=(dh, dw), mode="bicubic", align_corners=False)[0].permute(1, 2, 0).reshape(
            dh * dw, c).mul(g).sum().backward()
        assert _rel(gsrc.cpu(), s.grad) < 1e-5, (dh, dw)
    tab = torch.randn(15, 64)
    for rows in (31, 39, 15, 9):
        ref = F.interpolate(tab.t()[None], size=rows, mode="linear")[0].t()
        out = torch.empty(rows, 64, device=dev)
        ops.call("aldi_linear_resize_rows", tab.to(dev), 15, out, rows, 64, 0)
        assert _rel(out.cpu(), ref) < 1e-6, rows
        g = torch.randn(rows, 64)
        gt = torch.zeros(15, 64, device=dev)
        ops.call("aldi_linear_resize_rows", gt, 15, g.to(dev), rows, 64, 1)
        t = tab.clone().requires_grad_(True)
        F.interpolate(t.t()[None], size=rows, mode="linear")[0].t().mul(g).sum().backward()
        assert _rel(gt.cpu(), t.grad) < 1e-5, rows


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool2x2_and_broadcast_add(dtype):
    from aldi_b200 import lib as _l, ops
    dev = _dev()
    dtc = _l.BF16 if dtype == torch.bfloat16 else _l.F32
    n, h, w, c = 2, 8, 10, 64
    x = torch.randn(n, h, w, c).to(dtype)
    x[0, 0, 0, 0] = x[0, 0, 1, 0] = 9.0            # a tie: the first maximum takes the gradient
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    g = torch.randn_like(yr).to(dtype)
    yr.backward(g.float())
    y = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=dtype)
    ops.call("aldi_maxpool2x2", x.to(dev), y, dtc, n, h, w, c)
    assert torch.equal(y.float().cpu(), yr.detach().permute(0, 2, 3, 1))
    dx = torch.full((n, h, w, c), 3.0, device=dev, dtype=dtype)
    ops.call("aldi_maxpool2x2_backward", x.to(dev), g.permute(0, 2, 3, 1).contiguous().to(dev), dx, dtc, n, h, w, c)
    assert torch.equal(dx.float().cpu(), xr.grad.permute(0, 2, 3, 1))
    pos = torch.randn(h * w, c)
    xd = x.to(dev).clone()
    ops.call("aldi_add_rows_bcast", xd, pos.to(dev), n, h * w, c, dtc)
    assert _rel(xd.float().cpu(), (x.float() + pos.view(1, h, w, c)).to(dtype).float()) < 1e-6
    acc = torch.ones(h * w, c, device=dev)
    ops.call("aldi_sum_over_batch", x.to(dev), n, h * w, c, dtc, acc)
    assert _rel(acc.cpu(), 1 + x.float().sum(0).view(h * w, c)) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# attention op
# ---------------------------------------------------------------------------------------------------------------------
def _ref_attention(qkv, rel, gh, gw, heads, scale):
    """float64 restatement of vit.Attention's core with the relative-position products given as a tensor."""
    b, t, _ = qkv.shape
    q, k, v = qkv.view(b, t, 3, heads, 64).permute(2, 0, 3, 1, 4)                # (b, heads, t, 64)
    s = (q * scale) @ k.transpose(-2, -1)
    ys, xs = torch.arange(t) // gw, torch.arange(t) % gw
    ih = gh - 1 + ys[:, None] - ys[None, :]
    iw = 2 * gh - 1 + gw - 1 + xs[:, None] - xs[None, :]
    r = rel.permute(0, 2, 1, 3)                                                  # (b, heads, t, columns)
    s = s + r.gather(3, ih.expand(b, heads, t, t)) + r.gather(3, iw.expand(b, heads, t, t))
    lse = torch.logsumexp(s, -1)
    out = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(b, t, heads * 64)
    return out, lse


def _run_attention(qkv, rel, gh, gw, heads, scale, dout, impl):
    """rel / the returned drel are in the GEMM layout (b, tokens, heads, columns): the key-major arrays the kernels consume
    are made by aldi_relpos_transpose and the gradients brought back by its inverse, as aldi_b200/vit.py does."""
    from aldi_b200 import lib as _l, ops
    L = _l.load()
    b, t, _ = qkv.shape
    dev, dt = qkv.device, qkv.dtype
    out = torch.zeros(b, t, heads * 64, device=dev, dtype=dt)
    lse = torch.zeros(b, heads, t, device=dev)
    dqkv = torch.full_like(qkv, 5.0)
    rel_h, rel_w = torch.full((b, heads, gh, t), 7.0, device=dev), torch.full((b, heads, gw, t), 7.0, device=dev)
    ops.call("aldi_relpos_transpose", rel, rel.shape[3], rel_h, rel_w, b, gh, gw, heads, 0)
    drel_h, drel_w = torch.full_like(rel_h, 5.0), torch.full_like(rel_w, 5.0)
    delta = torch.zeros_like(lse)
    p = _l.AttnParams()
    p.qkv, p.batch, p.gh, p.gw, p.heads = qkv.data_ptr(), b, gh, gw, heads
    p.row_stride, p.batch_stride = qkv.stride(1), qkv.stride(0)
    p.rel_h, p.rel_w, p.scale = rel_h.data_ptr(), rel_w.data_ptr(), scale
    p.dtype = _l.BF16 if dt == torch.bfloat16 else _l.F32
    p.out, p.out_stride, p.out_batch_stride, p.lse = out.data_ptr(), out.stride(1), out.stride(0), lse.data_ptr()
    p.dout, p.dqkv, p.drel_h, p.drel_w, p.delta, p.impl = (dout.data_ptr(), dqkv.data_ptr(), drel_h.data_ptr(), drel_w.data_ptr(),
                                                         delta.data_ptr(), impl)
    _l.check(L.aldi_attention_forward(ctypes.byref(p), ops._stream()), "fwd")
    _l.check(L.aldi_attention_backward(ctypes.byref(p), ops._stream()), "bwd")
    drel = torch.full_like(rel, 5.0)
    ops.call("aldi_relpos_transpose", drel, rel.shape[3], drel_h, drel_w, b, gh, gw, heads, 1)
    torch.cuda.synchronize()
    return out, lse, dqkv, drel


ATTN_CASES = [(3, 14, 14, 2), (2, 8, 10, 2), (1, 16, 24, 3), (1, 5, 37, 1)]


def _attention_problem(b, gh, gw, heads, seed=0):
    g = torch.Generator().manual_seed(seed)
    t, nr = gh * gw, 2 * gh - 1 + 2 * gw - 1
    qkv = torch.randn(b, t, 3 * heads * 64, generator=g)
    tables = torch.randn(nr, 64, generator=g) * 0.2
    dout = torch.randn(b, t, heads * 64, generator=g)
    return qkv, tables, dout, (nr + 63) // 64 * 64


@pytest.mark.parametrize("b,gh,gw,heads", ATTN_CASES)
def test_attention_fp32_matches_reference(b, gh, gw, heads):
    dev = _dev()
    qkv, tables, dout, nrp = _attention_problem(b, gh, gw, heads)
    t = gh * gw
    scale = 0.125
    q = qkv.view(b, t, 3, heads, 64)[:, :, 0]
    rel = torch.zeros(b, t, heads, nrp)
    rel[..., :tables.shape[0]] = q @ tables.t()
    qd = qkv.double().requires_grad_(True)
    rd = rel.double().requires_grad_(True)
    out_r, lse_r = _ref_attention(qd, rd, gh, gw, heads, scale)
    out_r.backward(dout.double())
    out, lse, dqkv, drel = _run_attention(qkv.to(dev), rel.to(dev), gh, gw, heads, scale, dout.to(dev), 0)
    assert _rel(out.cpu(), out_r.detach()) < 2e-5
    assert _rel(lse.cpu(), lse_r.detach()) < 2e-5
    assert _rel(dqkv.cpu(), qd.grad) < 5e-5
    assert _rel(drel.cpu(), rd.grad) < 5e-5


@pytest.mark.parametrize("b,gh,gw,heads", ATTN_CASES + [(1, 64, 64, 1)])
def test_attention_tcgen05_matches_reference_and_cuda_core_kernel(b, gh, gw, heads):
    """The benchmarked bf16 kernels (UMMA tiles over 16 x 8-token patches) against float64 on the same bf16-rounded inputs,
    and against the CUDA-core kernel run on those inputs (impl = 1)."""
    dev = _dev()
    qkv, tables, dout, nrp = _attention_problem(b, gh, gw, heads, seed=1)
    t = gh * gw
    scale = 0.125
    qkv, dout = qkv.bfloat16(), dout.bfloat16()
    q = qkv.float().view(b, t, 3, heads, 64)[:, :, 0]
    rel = torch.zeros(b, t, heads, nrp)
    rel[..., :tables.shape[0]] = q @ tables.t()
    qd = qkv.double().requires_grad_(True)
    rd = rel.double().requires_grad_(True)
    out_r, lse_r = _ref_attention(qd, rd, gh, gw, heads, scale)
    out_r.backward(dout.double())
    res_tc = _run_attention(qkv.to(dev), rel.to(dev), gh, gw, heads, scale, dout.to(dev), 0)
    res_cc = _run_attention(qkv.to(dev), rel.to(dev), gh, gw, heads, scale, dout.to(dev), 1)
    refs = (out_r.detach(), lse_r.detach(), qd.grad, rd.grad)
    names = ("out", "lse", "dqkv", "drelpos")
    # P, dS, O and the gradients pass through bf16: 2^-8 relative on values, accumulated over the keys
    tol = {"out": 1.5e-2, "lse": 1e-4, "dqkv": 2.5e-2, "drelpos": 2e-2}
    for name, a, c, r in zip(names, res_tc, res_cc, refs):
        assert torch.isfinite(a.float()).all(), name
        assert _rel(c.cpu(), r) < tol[name], ("cuda-core", name, _rel(c.cpu(), r))
        assert _rel(a.cpu(), r) < tol[name], ("tcgen05", name, _rel(a.cpu(), r))


# ---------------------------------------------------------------------------------------------------------------------
# whole backbone
# ---------------------------------------------------------------------------------------------------------------------
SMALL = dict(embed_dim=128, depth=4, num_heads=2, drop_path_rate=0.2, window_block_indexes=(0, 2), lr_decay_rate=0.7)
MEAN, STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)


def _oracle_backbone(sd, img_size):
    from oracle import vit_ref
    net = vit_ref.ViT(img_size=img_size, embed_dim=SMALL["embed_dim"], depth=SMALL["depth"], num_heads=SMALL["num_heads"],
                      drop_path_rate=SMALL["drop_path_rate"], window_block_indexes=SMALL["window_block_indexes"])
    m = vit_ref.SimpleFeaturePyramid(net)
    m.load_state_dict(sd, strict=True)
    return m.double()


def _backbone_case(hw, seed=0):
    from aldi_b200 import vit
    layout = vit.ViTLayout(img_size=128, **SMALL)
    sd = vit.synthetic_state_dict(layout, seed=seed, rel_pos_std=0.3)
    g = torch.Generator().manual_seed(seed + 1)
    for k in sd:                                   # non-trivial norms / biases so their gradients are exercised
        if k.endswith(".bias") or (sd[k].dim() == 1):
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    img = torch.randint(0, 256, (2, 3, hw[0], hw[1]), generator=g, dtype=torch.uint8)
    return layout, sd, img, g


@pytest.mark.parametrize("hw", [(128, 160), (256, 320)])
def test_vitdet_backbone_forward_backward_matches_oracle(hw):
    """fp32 parity mode: pyramid outputs and EVERY parameter gradient against the oracle (float64), with DropPath factors,
    zero-padded edge windows (8 x 10 and 16 x 20 token grids under 14 x 14 windows), linearly resampled relative-position
    tables (global blocks on a non-square grid) and the bicubically resampled pos_embed."""
    from aldi_b200 import vit
    dev = _dev()
    layout, sd, img, g = _backbone_case(hw)
    net = vit.ViTDetBackbone(sd, size=None, dtype="fp32", device=dev, img_size=128, pixel_mean=MEAN, pixel_std=STD, **SMALL)
    keep = net.draw_keep_masks(2, generator=g)
    assert keep[0] is None and keep[1] is None and any(k is not None for k in keep)
    sizes = torch.tensor([[hw[0], hw[1]]] * 2, dtype=torch.int32, device=dev)
    outs = net.forward(img.to(dev), sizes, keep_masks=keep, save=True)
    ora = _oracle_backbone(sd, 128)
    ora.train()
    ora.net.keep_queue = [k.double() for k in keep if k is not None]
    x = (img.double() - torch.tensor(MEAN).view(1, 3, 1, 1)) / torch.tensor(STD).view(1, 3, 1, 1)
    ref = ora(x)
    assert not ora.net.keep_queue
    gs, loss = {}, 0.0
    for name in ("p2", "p3", "p4", "p5"):
        r = ref[name].permute(0, 2, 3, 1)
        assert _rel(outs[name].cpu(), r.detach()) < 2e-4, name
        gs[name] = torch.randn(r.shape, generator=g)
        loss = loss + (r * gs[name].double()).sum()
    loss.backward()
    net.backward({k: v.to(dev) for k, v in gs.items()})
    got = layout.unpack(net.grad)
    worst = {}
    for k, p in ora.named_parameters():
        worst[k] = _rel(got[k], p.grad)
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]


def test_vitdet_backbone_bf16_tracks_fp32():
    """The benchmarked arithmetic (bf16 activations, tcgen05 GEMMs + attention) against the fp32 parity mode on the same
    weights and image: bf16-level agreement of the pyramid and of the gradients' direction."""
    from aldi_b200 import vit
    dev = _dev()
    layout, sd, img, g = _backbone_case((256, 320), seed=3)
    sizes = torch.tensor([[256, 320]] * 2, dtype=torch.int32, device=dev)
    res = {}
    grads = {}
    for mode in ("fp32", "bf16"):
        net = vit.ViTDetBackbone(sd, size=None, dtype=mode, device=dev, img_size=128, pixel_mean=MEAN, pixel_std=STD, **SMALL)
        outs = net.forward(img.to(dev), sizes, keep_masks=None, save=True)
        if not grads:
            grads = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}
        net.backward({k: grads[k].to(dev, outs[k].dtype) for k in outs})
        res[mode] = ({k: v.float().cpu() for k, v in outs.items()}, net.grad.cpu().clone())
    for k in res["fp32"][0]:
        assert _rel(res["bf16"][0][k], res["fp32"][0][k]) < 6e-2, k
    a, b = res["bf16"][1].double(), res["fp32"][1].double()
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 0.995, cos
    assert abs(float(a.norm() / b.norm()) - 1) < 5e-2


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] shape
# ---------------------------------------------------------------------------------------------------------------------
FULL = dict(embed_dim=768, depth=2, num_heads=12, drop_path_rate=0.0, window_block_indexes=(0,), lr_decay_rate=1.0)


def test_vitdet_b_blocks_at_the_baseline_shape_match_oracle():
    """ViT-B widths (768 channels, 12 heads, hidden 3072) on a 1024 x 1024 image = the 64 x 64 token grid of BASELINE
    configs[2]: one windowed block (25 windows of 14 x 14 with zero-padded edges) and one global block (4096 tokens, the
    127-row relative-position tables at their native length), then the SimpleFeaturePyramid -- the tile variants, TMA patch
    geometry and launch sizes the benchmark runs.  fp32 parity mode against the float64 oracle (pyramid + every parameter
    gradient), and the benchmarked bf16 tcgen05 path against the fp32 mode."""
    from aldi_b200 import vit
    dev = _dev()
    layout = vit.ViTLayout(img_size=1024, **FULL)
    sd = vit.synthetic_state_dict(layout, seed=5, rel_pos_std=0.2)
    g = torch.Generator().manual_seed(6)
    for k in sd:
        if sd[k].dim() == 1:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    img = torch.randint(0, 256, (1, 3, 1024, 1024), generator=g, dtype=torch.uint8)
    sizes = torch.tensor([[1024, 1024]], dtype=torch.int32, device=dev)
    from oracle import vit_ref
    net = vit_ref.ViT(img_size=1024, embed_dim=768, depth=2, num_heads=12, drop_path_rate=0.0, window_block_indexes=(0,))
    ora = vit_ref.SimpleFeaturePyramid(net)
    ora.load_state_dict(sd, strict=True)
    ora = ora.double().train()
    x = (img.double() - torch.tensor(MEAN).view(1, 3, 1, 1)) / torch.tensor(STD).view(1, 3, 1, 1)
    ref = ora(x)
    gs, loss = {}, 0.0
    for name in ("p2", "p3", "p4", "p5"):
        r = ref[name].permute(0, 2, 3, 1)
        gs[name] = torch.randn(r.shape, generator=g)
        loss = loss + (r * gs[name].double()).sum()
    loss.backward()
    res = {}
    for mode in ("fp32", "bf16"):
        net_d = vit.ViTDetBackbone(sd, size=None, dtype=mode, device=dev, img_size=1024, pixel_mean=MEAN, pixel_std=STD, **FULL)
        outs = net_d.forward(img.to(dev), sizes, keep_masks=None, save=True)
        net_d.backward({k: v.to(dev, outs[k].dtype) for k, v in gs.items()})
        res[mode] = ({k: v.float().cpu() for k, v in outs.items()}, layout.unpack(net_d.grad))
        del net_d, outs
        torch.cuda.empty_cache()
    for name in ("p2", "p3", "p4", "p5"):
        r = ref[name].permute(0, 2, 3, 1).detach()
        assert _rel(res["fp32"][0][name], r) < 2e-4, name
        assert _rel(res["bf16"][0][name], r) < 6e-2, name
    bad = {}
    cos_num = cos_a = cos_b = 0.0
    for k, p in ora.named_parameters():
        e = _rel(res["fp32"][1][k], p.grad)
        if e > 1e-3:
            bad[k] = e
        a, b = res["bf16"][1][k].double().flatten(), p.grad.flatten()
        cos_num += float(a @ b); cos_a += float(a @ a); cos_b += float(b @ b)
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]
    assert cos_num / (cos_a * cos_b) ** 0.5 > 0.995, cos_num / (cos_a * cos_b) ** 0.5
