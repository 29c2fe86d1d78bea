"""GPU diagnostics for the conv kernels: each case runs in its own subprocess (a trap in one kernel must not
take the others down) and prints max-abs / max-rel error against a torch fp32 CPU reference.

usage: python tools/gpu_diag.py            (run all cases, log to gpurun_out/diag.log)
       python tools/gpu_diag.py --case X   (run one case in-process)
"""
import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _bf(x):
    import torch
    return x.to(torch.bfloat16)


def ref_conv(x_nhwc, w_ohwi, pad, stride=1):
    """fp32 CPU reference: channels-last in, channels-last out."""
    import torch.nn.functional as F
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_ohwi.float().permute(0, 3, 1, 2)
    return F.conv2d(x, w, padding=pad, stride=stride).permute(0, 2, 3, 1).contiguous()


def report(name, got, want, tol=2e-2):
    import torch
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs()
    denom = want.abs().max().item() + 1e-12
    rel = err.max().item() / denom
    bad = (err > tol * denom).float().mean().item()
    print("[%s] max_abs_err=%.4e ref_max=%.4e rel=%.4e frac_bad=%.4f %s" %
          (name, err.max().item(), denom, rel, bad, "OK" if rel < tol else "FAIL"), flush=True)
    if rel >= tol:
        # error structure: which rows (pixels) / channels are wrong
        e2 = err.reshape(-1, err.shape[-1])
        rows_bad = (e2.max(dim=1).values > tol * denom).nonzero().flatten()
        cols_bad = (e2.max(dim=0).values > tol * denom).nonzero().flatten()
        print("   bad rows: %d of %d, first %s" % (rows_bad.numel(), e2.shape[0], rows_bad[:16].tolist()))
        print("   bad cols: %d of %d, first %s" % (cols_bad.numel(), e2.shape[1], cols_bad[:16].tolist()))
        print("   got[0,:8]=%s\n   want[0,:8]=%s" % (got.reshape(-1, got.shape[-1])[0, :8].tolist(),
                                                      want.reshape(-1, want.shape[-1])[0, :8].tolist()))
    return rel < tol


def make_case(n, h, w, cin, cout, k, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(cout, k, k, cin, generator=g) / (k * k * cin) ** 0.5
    return _bf(x), _bf(wt)


def run_fwd(name, n, h, w, cin, cout, k, dtype="bf16", **epi):
    import torch
    from aldi_b200 import ops
    x, wt = make_case(n, h, w, cin, cout, k)
    pad = k // 2
    want = ref_conv(x, wt, pad)
    dev = "cuda"
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    xd = x.to(dev, tdt)
    cout_p = (cout + 63) // 64 * 64
    wp = torch.zeros(cout_p, k * k * cin, device=dev, dtype=tdt)
    ops.pack_weight(wt.float().to(dev).contiguous(), wp, cout=cout, taps=k * k, cin=cin, cout_p=cout_p, cin_p=cin)
    out_f32 = epi.pop("out_f32", dtype != "bf16")
    out = torch.full((n, h, w, cout), 7.0, device=dev, dtype=torch.float32 if out_f32 else tdt)
    kw = {}
    g = torch.Generator().manual_seed(1)
    if epi.get("scale"):
        sc = torch.rand(cout_p, generator=g) + 0.5
        bi = torch.randn(cout_p, generator=g)
        kw["scale"] = sc.to(dev); kw["bias"] = bi.to(dev)
        want = want * sc[:cout] + bi[:cout]
    if epi.get("res") == 1:
        r = _bf(torch.randn(n, h, w, cout, generator=g))
        kw["residual"] = r.to(dev, tdt); kw["res_mode"] = 1
        want = want + r.float()
    if epi.get("res") == 2:
        r = _bf(torch.randn(n, (h + 1) // 2, (w + 1) // 2, cout, generator=g))
        kw["residual"] = r.to(dev, tdt); kw["res_mode"] = 2
        up = r.float().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)[:, :h, :w]
        want = want + up
    if epi.get("relu"):
        kw["relu"] = True
        want = want.clamp_min(0)
    if epi.get("mask"):
        m = _bf(torch.randn(n, h, w, cout, generator=g))
        kw["mask"] = m.to(dev, tdt)
        want = want * (m.float() > 0)
    if epi.get("acc"):
        kw["accumulate"] = True
        base = _bf(torch.randn(n, h, w, cout, generator=g))
        out.copy_(base.to(dev))
        want = want + base.float()
    ops.conv(xd, wp, out, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad, **kw)
    torch.cuda.synchronize()
    return report(name, out, want)


def case_fwd_1x1_min():
    return run_fwd("fwd 1x1 64->64 1x8x16 (single tile, single k-block)", 1, 8, 16, 64, 64, 1)


def case_fwd_1x1_k256():
    return run_fwd("fwd 1x1 256->256 2x32x64", 2, 32, 64, 256, 256, 1)


def case_fwd_1x1_n128():
    return run_fwd("fwd 1x1 128->128 1x16x128", 1, 16, 128, 128, 128, 1)


def case_fwd_3x3():
    return run_fwd("fwd 3x3 64->128 1x20x24 (ragged tiles)", 1, 20, 24, 64, 128, 3)


def case_fwd_3x3_big():
    return run_fwd("fwd 3x3 256->256 2x64x96", 2, 64, 96, 256, 256, 3)


def case_fwd_epilogue():
    ok = run_fwd("fwd 3x3 epilogue scale+bias+res1+relu", 1, 16, 32, 64, 64, 3, scale=True, res=1, relu=True)
    ok &= run_fwd("fwd 1x1 epilogue bias+res2(upsample)", 2, 16, 32, 128, 256, 1, scale=True, res=2)
    ok &= run_fwd("fwd 1x1 epilogue mask+accumulate", 1, 16, 16, 64, 64, 1, mask=True, acc=True)
    ok &= run_fwd("fwd 1x1 f32 out, cout_store=15", 1, 16, 16, 256, 15, 1, out_f32=True, scale=True)
    return ok


def case_fwd_epilogue2():
    """TMA epilogue stress: many tiles per CTA, several N tiles, every operand combination the detector uses."""
    import torch
    from aldi_b200 import ops
    ok = run_fwd("fwd 1x1 64->256 +res1+relu, 4x96x160 (480 tiles)", 4, 96, 160, 64, 256, 1, scale=True, res=1, relu=True)
    ok &= run_fwd("fwd 1x1 128->512 +res1+relu, 2 N tiles", 2, 40, 72, 128, 512, 1, scale=True, res=1, relu=True)
    ok &= run_fwd("fwd 1x1 256->512 mask+res1 (conv1 dgrad join)", 2, 36, 52, 256, 512, 1, mask=True, res=1)
    ok &= run_fwd("fwd 3x3 128->128 mask", 2, 36, 52, 128, 128, 3, mask=True)
    ok &= run_fwd("fwd 1x1 512->256 +res2, ragged 50x70", 2, 50, 70, 512, 256, 1, scale=True, res=2)
    ok &= run_fwd("fwd 3x3 256->256 acc", 2, 30, 44, 256, 256, 3, acc=True)
    ok &= run_fwd("fwd 1x1 64->64 plain 3x7x9 (tiny ragged)", 3, 7, 9, 64, 64, 1)
    # accumulate into the even positions of a larger map (stride-2 dgrad scatter) with a mask view
    n, h, w, cin, cout = 2, 20, 28, 128, 256
    x, wt = make_case(n, h, w, cin, cout, 1)
    want = ref_conv(x, wt, 0)
    g = torch.Generator().manual_seed(11)
    big = _bf(torch.randn(n, 2 * h, 2 * w, cout, generator=g))
    m = _bf(torch.randn(n, 2 * h, 2 * w, cout, generator=g))
    wp = torch.zeros(cout, cin, device="cuda", dtype=torch.bfloat16)
    ops.pack_weight(wt.float().cuda().contiguous(), wp, cout=cout, taps=1, cin=cin, cout_p=cout, cin_p=cin)
    out = big.cuda().clone()
    ops.conv(x.cuda(), wp, out[:, ::2, ::2, :], mask=m.cuda()[:, ::2, ::2, :], accumulate=True)
    torch.cuda.synchronize()
    ref = big.float().clone()
    ref[:, ::2, ::2, :] += want * (m.float()[:, ::2, ::2, :] > 0)
    ok &= report("fwd 1x1 mask+acc into strided (::2, ::2) output view", out, ref)
    return ok


def case_stem():
    """bf16 stem: normalise + space-to-depth + 4x1-tap conv over the overlapping-row view vs F.conv2d(7x7, s2, p3)."""
    import torch
    import torch.nn.functional as F
    from aldi_b200 import arch
    from aldi_b200.detector import Detector, DetectorWeights, FlatLayout, PIXEL_MEAN
    ok = True
    for (n, h, w, vh, vw) in ((2, 64, 96, 64, 96), (2, 96, 160, 81, 150)):
        g = torch.Generator().manual_seed(h)
        img = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
        img[:, :, vh:, :] = 0
        img[:, :, :, vw:] = 0
        sizes = torch.tensor([[vh, vw]] * n, dtype=torch.int32)
        layout = FlatLayout(8)
        sd = arch.synthetic_state_dict(3)
        W = DetectorWeights(layout, layout.pack_state_dict(sd).cuda(), torch.bfloat16)
        W.refresh()
        det = Detector(8)
        # run only the stem part of backbone(): replicate its first lines
        from aldi_b200 import ops
        mean, std = ops.host_floats(PIXEL_MEAN), ops.host_floats((1.0, 1.0, 1.0))
        ho, wo = h // 2, w // 2
        wpad = wo + 4
        s2d = torch.empty(n, ho, wpad, 16, device="cuda", dtype=torch.bfloat16)
        ops.call("aldi_stem_s2d", img.cuda(), sizes.cuda(), s2d, n, h, w, mean, std)
        view = s2d.as_strided((n, ho, wo, 64), (ho * wpad * 16, wpad * 16, 16, 1))
        out = torch.empty(n, ho, wo, 64, device="cuda", dtype=torch.bfloat16)
        ops.conv(view, W.fwd["stem"], out, taps_h=4, taps_w=1, pad_h=2, pad_w=0, scale=W.scale["stem"],
                 bias=W.shift["stem"], relu=True)
        torch.cuda.synchronize()
        x = img.float() - torch.tensor(PIXEL_MEAN).view(1, 3, 1, 1)
        x[:, :, vh:, :] = 0
        x[:, :, :, vw:] = 0
        wt = sd["backbone.bottom_up.stem.conv1.weight"]
        y = F.conv2d(x.bfloat16().float(), wt.bfloat16().float(), stride=2, padding=3)
        bn = {k: sd["backbone.bottom_up.stem.conv1.norm." + k] for k in ("weight", "bias", "running_mean", "running_var")}
        sc = bn["weight"] * torch.rsqrt(bn["running_var"] + 1e-5)
        y = (y * sc.view(1, -1, 1, 1) + (bn["bias"] - bn["running_mean"] * sc).view(1, -1, 1, 1)).clamp_min(0)
        ok &= report("stem s2d %dx%dx%d (valid %dx%d)" % (n, h, w, vh, vw), out, y.permute(0, 2, 3, 1))
    return ok


def case_fwd_f32():
    ok = run_fwd("f32 fwd 3x3 64->128 1x20x24", 1, 20, 24, 64, 128, 3, dtype="f32")
    ok &= run_fwd("f32 fwd 3x3 epilogue", 1, 16, 32, 48, 40, 3, dtype="f32", scale=True, res=1, relu=True, mask=True)
    ok &= run_fwd("f32 fwd 1x1 res2 + acc", 2, 16, 32, 32, 64, 1, dtype="f32", res=2, acc=True)
    return ok


def case_stride2_view():
    import torch
    from aldi_b200 import ops
    ok = True
    for dtype in ("bf16", "f32"):
        n, h, w, cin, cout = 2, 32, 64, 128, 256
        x, wt = make_case(n, h, w, cin, cout, 1)
        want = ref_conv(x, wt, 0, stride=2)
        tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
        xd = x.to("cuda", tdt)
        wp = torch.zeros(cout, cin, device="cuda", dtype=tdt)
        ops.pack_weight(wt.float().cuda().contiguous(), wp, cout=cout, taps=1, cin=cin, cout_p=cout, cin_p=cin)
        out = torch.zeros(n, h // 2, w // 2, cout, device="cuda", dtype=tdt)
        ops.conv(xd[:, ::2, ::2, :], wp, out)
        torch.cuda.synchronize()
        ok &= report("%s fwd 1x1 stride-2 via strided view" % dtype, out, want)
    return ok


def _autograd_ref(x, wt, dy, pad, stride=1):
    import torch
    import torch.nn.functional as F
    xx = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ww = wt.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    y = F.conv2d(xx, ww, padding=pad, stride=stride)
    y.backward(dy.float().permute(0, 3, 1, 2))
    return xx.grad.permute(0, 2, 3, 1).contiguous(), ww.grad.permute(0, 2, 3, 1).contiguous()


def run_bwd(name, n, h, w, cin, cout, k, dtype="bf16"):
    import torch
    from aldi_b200 import ops
    x, wt = make_case(n, h, w, cin, cout, k)
    pad = k // 2
    g = torch.Generator().manual_seed(5)
    dy = _bf(torch.randn(n, h, w, cout, generator=g))
    sc = torch.rand(cout, generator=g) + 0.5
    dx_ref, dw_ref = _autograd_ref(x, wt, dy * sc, pad)   # y = conv*scale  =>  both grads see dy*scale
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    dev = "cuda"
    cin_p = (cin + 63) // 64 * 64
    cout_p = (cout + 63) // 64 * 64
    # data gradient = conv of dy with the flipped/transposed operand
    wd = torch.zeros(cin_p, k * k * cout_p, device=dev, dtype=tdt)
    ops.pack_weight(wt.float().to(dev).contiguous(), wd, dgrad=True, scale=sc.to(dev), cout=cout, taps=k * k, cin=cin,
                    cout_p=cout_p, cin_p=cin_p)
    dyd = torch.zeros(n, h, w, cout_p, device=dev, dtype=tdt)
    dyd[..., :cout] = dy.to(dev, tdt)
    dx = torch.zeros(n, h, w, cin, device=dev, dtype=tdt)
    ops.conv(dyd, wd, dx, taps_h=k, taps_w=k, pad_h=k - 1 - pad, pad_w=k - 1 - pad)
    xd = torch.zeros(n, h, w, cin_p, device=dev, dtype=tdt)
    xd[..., :cin] = x.to(dev, tdt)
    dw = torch.zeros(cout, k, k, cin, device=dev, dtype=torch.float32)
    ops.wgrad(xd, dyd, dw, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad, scale=torch.cat([sc, torch.ones(cout_p - cout)]).to(dev),
              cout_store=cout, cin_store=cin)
    torch.cuda.synchronize()
    ok = report(name + " dgrad", dx, dx_ref)
    ok &= report(name + " wgrad", dw, dw_ref)
    return ok


def case_bwd_1x1():
    return run_bwd("bwd 1x1 64->64 1x8x16", 1, 8, 16, 64, 64, 1)


def case_bwd_1x1_big():
    return run_bwd("bwd 1x1 256->128 2x32x64", 2, 32, 64, 256, 128, 1)


def case_bwd_3x3():
    return run_bwd("bwd 3x3 64->128 2x20x24", 2, 20, 24, 64, 128, 3)


def case_bwd_3x3_big():
    return run_bwd("bwd 3x3 256->256 2x32x48", 2, 32, 48, 256, 256, 3)


def case_bwd_mh2():
    """wgrad with the 256-row M tile at every BLOCK_N (cin 64 / 128 / 256) and ragged pixel counts."""
    os.environ["ALDI_WGRAD_MH2"] = "1"   # read once, at the first aldi_wgrad_tc call of this process
    ok = run_bwd("bwd 1x1 64->256 2x24x40 (N=64, M=256)", 2, 24, 40, 64, 256, 1)
    ok &= run_bwd("bwd 1x1 128->512 2x24x40 (N=128, M=256)", 2, 24, 40, 128, 512, 1)
    ok &= run_bwd("bwd 3x3 128->256 1x20x28 (N=128, M=256)", 1, 20, 28, 128, 256, 3)
    ok &= run_bwd("bwd 1x1 512->256 3x17x23 (N=256, M=256)", 3, 17, 23, 512, 256, 1)
    ok &= run_bwd("bwd 1x1 256->1024 2x64x64 (N=256, M=256, 4 M tiles)", 2, 64, 64, 256, 1024, 1)
    return ok


def case_bwd_f32():
    ok = run_bwd("f32 bwd 3x3 48->40 2x12x20", 2, 12, 20, 48, 40, 3, dtype="f32")
    ok &= run_bwd("f32 bwd 1x1 64->128 1x16x16", 1, 16, 16, 64, 128, 1, dtype="f32")
    return ok


def case_fc():
    """linear layer as a 1x1 conv over a (1,1,M,K) view."""
    import torch
    from aldi_b200 import ops
    M, K, N = 300, 1024, 1024
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(M, K, generator=g))
    w = _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    want = x.float() @ w.float().t()
    xd = x.cuda().view(1, 1, M, K)
    wp = w.cuda().contiguous()
    out = torch.zeros(1, 1, M, N, device="cuda", dtype=torch.bfloat16)
    ops.conv(xd, wp, out)
    torch.cuda.synchronize()
    return report("fc 300x1024x1024", out.view(M, N), want)


def case_optim():
    import torch
    from aldi_b200 import ops
    g = torch.Generator().manual_seed(0)
    n = 1000003
    s = torch.randn(n, generator=g)
    t = torch.randn(n, generator=g)
    alpha = 0.9996
    want = s * (1 - alpha) + t * alpha
    td = t.cuda()
    ops.ema_update(td, s.cuda(), alpha)
    torch.cuda.synchronize()
    exact = torch.equal(td.cpu(), want)
    print("[ema] bit-exact vs torch expression: %s (max diff %.3e)" % (exact, (td.cpu() - want).abs().max().item()), flush=True)
    p = torch.randn(n, generator=g); m = torch.randn(n, generator=g); gr = torch.randn(n, generator=g)
    pp = torch.nn.Parameter(p.clone()); opt = torch.optim.SGD([pp], lr=0.06, momentum=0.9, weight_decay=1e-4)
    opt.state[pp]["momentum_buffer"] = m.clone(); pp.grad = gr.clone(); opt.step()
    pd, md = p.cuda(), m.cuda()
    ops.sgd_momentum_step(pd, md, gr.cuda(), 0.06, 1e-4, 0.9)
    torch.cuda.synchronize()
    d = (pd.cpu() - pp.detach()).abs().max().item()
    print("[sgd] max diff vs torch.optim.SGD: %.3e %s" % (d, "OK" if d < 1e-6 else "FAIL"), flush=True)
    return exact and d < 1e-6


def case_perf():
    """device-time the production shapes (bf16 tcgen05 path) and cross-check against the fp32 CUDA-core kernel."""
    import torch
    from aldi_b200 import ops
    shapes = [
        ("fpn_out p2 3x3 256->256 4x256x512", 4, 256, 512, 256, 256, 3),
        ("res3 3x3 128->128 4x128x256", 4, 128, 256, 128, 128, 3),
        ("res4 1x1 1024->256 4x64x128", 4, 64, 128, 1024, 256, 1),
        ("res4 1x1 256->1024 4x64x128", 4, 64, 128, 256, 1024, 1),
        ("res5 3x3 512->512 4x32x64", 4, 32, 64, 512, 512, 3),
    ]
    ok = True
    for name, n, h, w, cin, cout, k in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, k * k * cin, device="cuda", generator=g) / (k * k * cin) ** 0.5).bfloat16()
        out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        pad = k // 2
        for _ in range(3):
            ops.conv(x, wt, out, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.conv(x, wt, out, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * n * h * w * cin * cout * k * k
        print("[perf fwd] %s: %.3f ms  %.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)
        # check vs f32 kernel on a sub-sampled set of outputs
        ref = torch.empty(n, h, w, cout, device="cuda", dtype=torch.float32)
        ops.conv(x.float(), wt.float(), ref, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        torch.cuda.synchronize()
        ok &= report("perf-check " + name, out, ref)
        # wgrad timing
        dy = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
        dw = torch.zeros(cout, k * k * cin, device="cuda")
        for _ in range(2):
            ops.wgrad(x, dy, dw, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        dw.zero_()
        e0.record()
        for _ in range(reps):
            ops.wgrad(x, dy, dw, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("[perf wgrad] %s: %.3f ms  %.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)
        dwr = torch.zeros(cout, k * k * cin, device="cuda")
        ops.wgrad(x.float(), dy.float(), dwr, taps_h=k, taps_w=k, pad_h=pad, pad_w=pad)
        torch.cuda.synchronize()
        ok &= report("perf-check wgrad " + name, dw / reps, dwr)
    return ok


def case_perf_mem():
    """the HBM-bound ResNet layers at full size (for ncu): 1x1 expand + FrozenBN + residual + ReLU, and its dgrad."""
    import torch
    from aldi_b200 import ops
    n, h, w, cin, cout = 4, 256, 512, 64, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, device="cuda", generator=g) / cin ** 0.5).bfloat16()
    res = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    bi = torch.randn(cout, device="cuda", generator=g)
    out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, kw in (("fwd 1x1 64->256 +bn+res1+relu", dict(scale=sc, bias=bi, residual=res, res_mode=1, relu=True)),
                     ("fwd 1x1 64->256 plain", dict()),
                     ("dgrad-like 1x1 64->256 mask", dict(mask=res))):
        for _ in range(3):
            ops.conv(x, wt, out, **kw)
        e0.record()
        for _ in range(10):
            ops.conv(x, wt, out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = (x.numel() + out.numel() * (2 if kw else 1)) * 2
        print("[perf_mem] %s: %.3f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6), flush=True)
    return True


def case_perf_small():
    """the layers whose time is not explained by flops or HBM bytes (3x3 with few channels, small maps)."""
    import torch
    from aldi_b200 import ops
    shapes = [("3x3 64->64 4x256x512", 4, 256, 512, 64, 64, 3), ("3x3 128->128 4x128x256", 4, 128, 256, 128, 128, 3),
              ("3x3 256->256 4x64x128", 4, 64, 128, 256, 256, 3), ("1x1 256->1024 4x64x128", 4, 64, 128, 256, 1024, 1)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, n, h, w, cin, cout, k in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, k * k * cin, device="cuda", generator=g) / (k * k * cin) ** 0.5).bfloat16()
        out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.conv(x, wt, out, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2)
        e0.record()
        for _ in range(10):
            ops.conv(x, wt, out, taps_h=k, taps_w=k, pad_h=k // 2, pad_w=k // 2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tiles = n * h * w / 128 * max(1, cout // 256)
        print("[perf_small] %s: %.1f us  %.0f TFLOP/s  %.2f us/tile/SM" % (
            name, ms * 1e3, 2.0 * n * h * w * cin * cout * k * k / ms / 1e9, ms * 1e3 / (tiles / 148)), flush=True)
    return True


def case_perf_res4():
    """res4 conv3 (1x1 256->1024 + FrozenBN + residual + ReLU at 4x64x128) and conv1 (1x1 1024->256): the short
    launches that sit far above both rooflines in the step profile; run under ncu / with the ALDI_CONV_* knobs."""
    import torch
    from aldi_b200 import ops
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, n, h, w, cin, cout, with_res in (("conv3 1x1 256->1024 +res", 4, 64, 128, 256, 1024, True),
                                                ("conv1 1x1 1024->256", 4, 64, 128, 1024, 256, False),
                                                ("res5 conv3 1x1 512->2048 +res", 4, 32, 64, 512, 2048, True)):
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, cin, device="cuda", generator=g) / cin ** 0.5).bfloat16()
        res = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
        sc = torch.rand(cout, device="cuda", generator=g) + 0.5
        bi = torch.randn(cout, device="cuda", generator=g)
        out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        kw = dict(scale=sc, bias=bi, relu=True)
        if with_res:
            kw.update(residual=res, res_mode=1)
        ts = []
        for i in range(8):
            flush.zero_()
            e0.record()
            ops.conv(x, wt, out, **kw)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        nbytes = (x.numel() + out.numel() * (2 if with_res else 1) + wt.numel()) * 2
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        print("[perf_res4] %s: %.1f us (cold L2)  %.0f GB/s  %.0f TFLOP/s" % (
            name, t, nbytes / t / 1e3, 2.0 * n * h * w * cin * cout / t / 1e6), flush=True)
    return True


def case_halo():
    """3x3 convs with <= 128 input channels through the halo-tile path (ALDI_CONV_HALO=1|2 in the environment):
    numerics against the fp32 CUDA-core kernel on odd sizes, then timing at the production sizes."""
    import torch
    from aldi_b200 import ops
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, n, h, w, cin, cout, extra in (("c64 plain", 2, 48, 72, 64, 64, {}), ("c64 bn+relu", 1, 37, 50, 64, 64, {"bn": True}),
                                            ("c128 plain", 2, 32, 40, 128, 128, {}), ("c128 mask", 1, 40, 24, 128, 128, {"mask": True}),
                                            ("c128 acc", 1, 16, 16, 128, 128, {"acc": True})):
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, 9 * cin, device="cuda", generator=g) / (9 * cin) ** 0.5).bfloat16()
        kw, kwf = {}, {}
        if extra.get("bn"):
            sc = torch.rand(cout, device="cuda", generator=g) + 0.5
            bi = torch.randn(cout, device="cuda", generator=g)
            kw = kwf = dict(scale=sc, bias=bi, relu=True)
        out = torch.zeros(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        ref = torch.zeros(n, h, w, cout, device="cuda", dtype=torch.float32)
        if extra.get("mask"):
            m = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
            kw, kwf = dict(mask=m), dict(mask=m.float())
        if extra.get("acc"):
            out = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
            ref = out.float().clone()
            kw = kwf = dict(accumulate=True)
        ops.conv(x, wt, out, taps_h=3, taps_w=3, pad_h=1, pad_w=1, **kw)
        ops.conv(x.float(), wt.float(), ref, taps_h=3, taps_w=3, pad_h=1, pad_w=1, **kwf)
        torch.cuda.synchronize()
        ok &= report("halo " + name, out, ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, n, h, w, c in (("3x3 64->64 4x256x512", 4, 256, 512, 64), ("3x3 64->64 8x256x512", 8, 256, 512, 64),
                             ("3x3 128->128 8x128x256", 8, 128, 256, 128)):
        x = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(c, 9 * c, device="cuda", generator=g) / (9 * c) ** 0.5).bfloat16()
        out = torch.empty(n, h, w, c, device="cuda", dtype=torch.bfloat16)
        ts = []
        for _ in range(8):
            flush.zero_()
            e0.record()
            ops.conv(x, wt, out, taps_h=3, taps_w=3, pad_h=1, pad_w=1)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts[2:])[3]
        print("[halo perf] %s: %.1f us  %.0f TFLOP/s" % (name, t, 2.0 * n * h * w * 9 * c * c / t / 1e6), flush=True)
    return ok


CASES = [
    "optim", "fwd_f32", "bwd_f32",
    "fwd_1x1_min", "fwd_1x1_k256", "fwd_1x1_n128", "fwd_3x3", "fwd_3x3_big", "fwd_epilogue", "fwd_epilogue2",
    "stride2_view", "fc", "stem",
    "bwd_1x1", "bwd_1x1_big", "bwd_3x3", "bwd_3x3_big", "bwd_mh2", "perf",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    if args.case:
        ok = globals()["case_" + args.case]()
        sys.exit(0 if ok else 1)
    os.makedirs("gpurun_out", exist_ok=True)
    log = open("gpurun_out/diag.log", "w")
    summary = []
    for c in (args.only or CASES):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True,
                               timeout=300)
            out, rc = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = "TIMEOUT\n" + str(e.stdout or "") + str(e.stderr or ""), -9
        msg = "=== case %s rc=%d (%.1fs)\n%s\n" % (c, rc, time.time() - t0, out[-6000:])
        print(msg, flush=True)
        log.write(msg); log.flush()
        summary.append((c, rc))
    s = "SUMMARY: " + " ".join("%s=%s" % (c, "ok" if rc == 0 else "FAIL(%d)" % rc) for c, rc in summary)
    print(s)
    log.write(s + "\n")


if __name__ == "__main__":
    main()
