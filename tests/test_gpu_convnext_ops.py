"""csrc/convnext.cu building blocks against the torch functions the reference's ConvNeXt calls on its device
(aldi/backbone.py:189-346: F.layer_norm / the channels_first formula, nn.Conv2d(groups=dim), nn.GELU, layer scale +
DropPath + residual, the stride-k patchify convs) and torch.optim.AdamW.  fp32 parity mode at 1e-4, bf16 at bf16
resolution."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DT = {"fp32": (torch.float32, 0, 2e-4), "bf16": (torch.bfloat16, 1, 2e-2)}


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("c,stride", [(96, 128), (192, 192), (1536, 1536), (40, 64)])
def test_layernorm(mode, c, stride):
    from aldi_b200 import ops
    dt, code, tol = DT[mode]
    g = torch.Generator().manual_seed(c)
    rows = 3 * 17 * 5
    x = torch.randn(rows, stride, generator=g) * 2 + 0.5
    x[:, c:] = 0
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    dy = torch.randn(rows, stride, generator=g)
    dy[:, c:] = 0
    xq, dyq = x.to(dt), dy.to(dt)
    xr = xq.float()[:, :c].clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (c,), gr, br, 1e-6)
    yr.backward(dyq.float()[:, :c])
    xd, dyd = xq.cuda(), dyq.cuda()
    y = torch.empty_like(xd)
    stats = torch.empty(rows, 2, device="cuda")
    ops.call("aldi_layernorm_forward", xd, gamma.cuda(), beta.cuda(), 1e-6, rows, c, stride, code, y, stats)
    dx = torch.empty_like(xd)
    dgam, dbet = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    ops.call("aldi_layernorm_backward", xd, gamma.cuda(), stats, dyd, rows, c, stride, code, dx, 0, dgam, dbet)
    torch.cuda.synchronize()
    assert rel(y.cpu().float()[:, :c], yr.detach()) < tol
    assert float(y[:, c:].abs().max()) == 0.0 if stride > c else True
    assert rel(dx.cpu().float()[:, :c], xr.grad) < tol
    assert rel(dgam.cpu(), gr.grad) < tol and rel(dbet.cpu(), br.grad) < tol


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("c,stride,h,w", [(96, 128, 13, 19), (192, 192, 8, 8), (24, 64, 5, 30), (160, 192, 37, 50)])
def test_dwconv7(mode, c, stride, h, w):
    from aldi_b200 import ops
    dt, code, tol = DT[mode]
    g = torch.Generator().manual_seed(h)
    n = 2
    x = torch.randn(n, h, w, stride, generator=g)
    x[..., c:] = 0
    wt, b = torch.randn(c, 1, 7, 7, generator=g) * 0.1, torch.randn(c, generator=g)
    dy = torch.randn(n, h, w, stride, generator=g)
    dy[..., c:] = 0
    xq, dyq = x.to(dt), dy.to(dt)
    xr = xq.float()[..., :c].permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr, brr = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, brr, padding=3, groups=c)
    yr.backward(dyq.float()[..., :c].permute(0, 3, 1, 2))
    xd, dyd, wd = xq.cuda(), dyq.cuda(), wt.reshape(c, 49).contiguous().cuda()
    y = torch.zeros_like(xd)
    ops.call("aldi_dwconv7", xd, wd, b.cuda(), n, h, w, c, stride, code, 0, y, 0)
    dx = torch.zeros_like(xd)
    ops.call("aldi_dwconv7", dyd, wd, None, n, h, w, c, stride, code, 1, dx, 0)
    dw = torch.zeros(c, 49, device="cuda")
    ops.call("aldi_dwconv7_wgrad", xd, dyd, n, h, w, c, stride, code, dw)
    torch.cuda.synchronize()
    assert rel(y.cpu().float()[..., :c], yr.detach().permute(0, 2, 3, 1)) < tol
    assert rel(dx.cpu().float()[..., :c], xr.grad.permute(0, 2, 3, 1)) < tol
    assert rel(dw.cpu(), wr.grad.reshape(c, 49)) < tol
    # accumulate: the data gradient is added onto what the residual branch already wrote; dw adds onto earlier passes
    base = torch.randn(n, h, w, stride, generator=g).to(dt)
    dx2 = base.cuda()
    ops.call("aldi_dwconv7", dyd, wd, None, n, h, w, c, stride, code, 1, dx2, 1)
    ops.call("aldi_dwconv7_wgrad", xd, dyd, n, h, w, c, stride, code, dw)
    torch.cuda.synchronize()
    assert rel(dx2.cpu().float()[..., :c], base.float()[..., :c] + xr.grad.permute(0, 2, 3, 1)) < tol
    assert rel(dw.cpu(), 2 * wr.grad.reshape(c, 49)) < tol
    if stride > c:                                                           # pad channels: old value + 0
        assert torch.equal(dx2.cpu()[..., c:], base[..., c:])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_gelu_and_layerscale(mode):
    from aldi_b200 import ops
    dt, code, tol = DT[mode]
    g = torch.Generator().manual_seed(0)
    n, hw, c, stride = 3, 35, 96, 128
    rows = n * hw
    h = (torch.randn(rows, 4 * stride, generator=g) * 2).to(dt)
    da = torch.randn(rows, 4 * stride, generator=g).to(dt)
    hr = h.float().clone().requires_grad_(True)
    ar = F.gelu(hr)
    ar.backward(da.float())
    a, dh = torch.empty_like(h, device="cuda"), torch.empty_like(h, device="cuda")
    ops.call("aldi_gelu", h.cuda(), None, a, h.numel(), code)
    ops.call("aldi_gelu", h.cuda(), da.cuda(), dh, h.numel(), code)
    torch.cuda.synchronize()
    assert rel(a.cpu().float(), ar.detach()) < tol and rel(dh.cpu().float(), hr.grad) < tol
    # x = input + drop_path(gamma * u)
    u = torch.randn(rows, stride, generator=g).to(dt)
    inp = torch.randn(rows, stride, generator=g).to(dt)
    dy = torch.randn(rows, stride, generator=g).to(dt)
    gamma = torch.randn(c, generator=g)
    keep = torch.tensor([1 / 0.8, 0.0, 1 / 0.8])
    ur, gr = u.float()[:, :c].clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    outr = inp.float()[:, :c] + (gr * ur) * keep.repeat_interleave(hw)[:, None]
    outr.backward(dy.float()[:, :c])
    out, du = torch.empty_like(u, device="cuda"), torch.empty_like(u, device="cuda")
    dg = torch.zeros(c, device="cuda")
    ops.call("aldi_layerscale_forward", u.cuda(), inp.cuda(), gamma.cuda(), keep.cuda(), rows, hw, c, stride, code, out)
    ops.call("aldi_layerscale_backward", u.cuda(), dy.cuda(), gamma.cuda(), keep.cuda(), rows, hw, c, stride, code, du, dg)
    torch.cuda.synchronize()
    assert rel(out.cpu().float()[:, :c], outr.detach()) < tol
    assert rel(du.cpu().float()[:, :c], ur.grad) < tol and rel(dg.cpu(), gr.grad) < tol


def test_space_to_depth_patchify_and_adamw():
    from aldi_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, h, w, b = 2, 12, 20, 2
    # c = 24: the 8-channel vector kernel (fp32 and bf16); c = 20: the per-element kernel
    for c, dt, code in [(24, torch.float32, 0), (24, torch.bfloat16, 1), (20, torch.float32, 0)]:
        x = torch.randn(n, h, w, 32, generator=g).to(dt)
        x[..., c:] = 0
        out = torch.full((n, h // b, w // b, 128), 7.0, device="cuda", dtype=dt)
        ops.call("aldi_space_to_depth", x.cuda(), out, n, h // b, w // b, b, c, 32, 128, code, 0)
        want = x[..., :c].reshape(n, h // b, b, w // b, b, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // b, w // b, b * b * c)
        assert torch.equal(out.cpu()[..., :b * b * c], want) and float(out[..., b * b * c:].float().abs().max()) == 0.0
        back = torch.zeros(n, h, w, 32, device="cuda", dtype=dt)
        ops.call("aldi_space_to_depth", out, back, n, h // b, w // b, b, c, 32, 128, code, 1)
        assert torch.equal(back.cpu()[..., :c], x[..., :c])
    # stem patches: conv2d(k=4, s=4) == patches @ weight(OHWI)
    img = torch.randint(0, 256, (2, 3, 16, 24), generator=g, dtype=torch.uint8)
    sizes = torch.tensor([[16, 24], [13, 21]], dtype=torch.int32)
    mean, std = (103.53, 116.28, 123.675), (57.0, 57.1, 57.4)
    patches = torch.empty(2, 4, 6, 64, device="cuda")
    ops.call("aldi_patchify_image", img.cuda(), sizes.cuda(), patches, 2, 16, 24, 4, 64, 0, ops.host_floats(mean), ops.host_floats(std))
    xf = (img.float() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    xf[1, :, 13:, :] = 0
    xf[1, :, :, 21:] = 0
    wt = torch.randn(8, 3, 4, 4, generator=g)
    want = F.conv2d(xf, wt, stride=4).permute(0, 2, 3, 1)
    got = patches.cpu()[..., :48] @ wt.permute(0, 2, 3, 1).reshape(8, 48).t()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4)
    # AdamW
    p = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(3)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    pd, m, v = p.cuda(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.call("aldi_adamw_step", pd, m, v, gr.cuda(), 1000, 1e-3, 0.9, 0.999, 1e-8, 0.05, i + 1, 1.0)
    torch.cuda.synchronize()
    assert torch.allclose(pd.cpu(), pr.detach(), rtol=1e-5, atol=1e-6)
