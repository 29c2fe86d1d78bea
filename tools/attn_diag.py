"""Diagnosis of the tcgen05 attention kernels on a GPU box: forward and backward of a few problem sizes against float64 on the
same bf16 inputs, with the error broken down by output (out / lse / dq / dk / dv / drelpos-h / drelpos-w) and by query tile,
so a wrong descriptor, swizzle or mask shows up as a pattern.  `python tools/attn_diag.py [fwd|bwd|all] [impl]`."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

from aldi_b200 import lib as _l, ops  # noqa: E402
import test_gpu_vit as tv  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def tile_errors(a, b, gh, gw):
    """max abs error per 16 x 8 token patch (rows of the printed grid = patch rows)."""
    e = (a.double() - b.double()).abs().amax(dim=-1)          # (batch, tokens)
    e = e.amax(0).view(gh, gw)
    th, tw = (gh + 7) // 8, (gw + 15) // 16
    out = []
    for i in range(th):
        out.append(["%.2e" % float(e[i * 8:(i + 1) * 8, j * 16:(j + 1) * 16].max()) for j in range(tw)])
    return out


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    impl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda:0")
    L = _l.load()
    cases = [(1, 8, 16, 1), (3, 14, 14, 2), (2, 8, 10, 2), (1, 16, 24, 3), (1, 5, 37, 1), (1, 64, 64, 1)]
    for b, gh, gw, heads in cases:
        qkv, tables, dout, nrp = tv._attention_problem(b, gh, gw, heads, seed=1)
        t, scale = gh * gw, 0.125
        qkv, dout = qkv.bfloat16(), dout.bfloat16()
        q = qkv.float().view(b, t, 3, heads, 64)[:, :, 0]
        relp = torch.zeros(b, t, heads, nrp)
        relp[..., :tables.shape[0]] = q @ tables.t()
        qd = qkv.double().requires_grad_(True)
        rd = relp.double().requires_grad_(True)
        out_r, lse_r = tv._ref_attention(qd, rd, gh, gw, heads, scale)
        out_r.backward(dout.double())
        qkv_d, rel_d, dout_d = qkv.to(dev), relp.to(dev), dout.to(dev)
        out = torch.zeros(b, t, heads * 64, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(b, heads, t, device=dev)
        dqkv = torch.full_like(qkv_d, 5.0)
        drel = torch.full_like(rel_d, 5.0)
        rel_h, rel_w = torch.zeros(b, heads, gh, t, device=dev), torch.zeros(b, heads, gw, t, device=dev)
        ops.call("aldi_relpos_transpose", rel_d, nrp, rel_h, rel_w, b, gh, gw, heads, 0)
        drel_h, drel_w = torch.full_like(rel_h, 5.0), torch.full_like(rel_w, 5.0)
        delta = torch.zeros_like(lse)
        p = _l.AttnParams()
        p.qkv, p.batch, p.gh, p.gw, p.heads = qkv_d.data_ptr(), b, gh, gw, heads
        p.row_stride, p.batch_stride = qkv_d.stride(1), qkv_d.stride(0)
        p.rel_h, p.rel_w, p.scale, p.dtype = rel_h.data_ptr(), rel_w.data_ptr(), scale, _l.BF16
        p.out, p.out_stride, p.out_batch_stride, p.lse = out.data_ptr(), out.stride(1), out.stride(0), lse.data_ptr()
        p.dout, p.dqkv, p.drel_h, p.drel_w, p.delta, p.impl = (dout_d.data_ptr(), dqkv.data_ptr(), drel_h.data_ptr(),
                                                             drel_w.data_ptr(), delta.data_ptr(), impl)
        print("case b=%d grid=%dx%d heads=%d impl=%d" % (b, gh, gw, heads, impl), flush=True)
        _l.check(L.aldi_attention_forward(ctypes.byref(p), ops._stream()), "fwd")
        torch.cuda.synchronize()
        print("  out %.3e  lse %.3e" % (rel(out.cpu(), out_r.detach()), rel(lse.cpu(), lse_r.detach())), flush=True)
        if rel(out.cpu(), out_r.detach()) > 2e-2:
            print("  out error per query patch:", tile_errors(out.cpu(), out_r.detach(), gh, gw))
            e = (out.cpu().double() - out_r.detach()).abs().view(b, t, heads, 64)
            print("  per head:", ["%.2e" % float(e[:, :, h].max()) for h in range(heads)],
                  " per 8-channel group:", ["%.2e" % float(e[..., c * 8:(c + 1) * 8].max()) for c in range(8)])
        if what == "fwd":
            continue
        # backward from the REFERENCE forward results so that a forward bug does not leak into this check
        out.copy_(out_r.detach().to(dev))
        lse.copy_(lse_r.detach().to(dev))
        _l.check(L.aldi_attention_backward(ctypes.byref(p), ops._stream()), "bwd")
        ops.call("aldi_relpos_transpose", drel, nrp, drel_h, drel_w, b, gh, gw, heads, 1)
        torch.cuda.synchronize()
        dim = heads * 64
        g, gr = dqkv.cpu().float(), qd.grad
        names = ("dq", "dk", "dv")
        for i, nm in enumerate(names):
            a, r = g[..., i * dim:(i + 1) * dim], gr[..., i * dim:(i + 1) * dim]
            print("  %s %.3e" % (nm, rel(a, r)), end="")
            if rel(a, r) > 3e-2:
                print("  per token patch:", tile_errors(a, r, gh, gw), end="")
            print(flush=True)
        nh = 2 * gh - 1
        dr, drr = drel.cpu(), rd.grad
        print("  drelpos-h %.3e  drelpos-w %.3e  pad-columns max %.1e" % (
            rel(dr[..., :nh], drr[..., :nh]), rel(dr[..., nh:nh + 2 * gw - 1], drr[..., nh:nh + 2 * gw - 1]),
            float(dr[..., nh + 2 * gw - 1:].abs().max()) if nh + 2 * gw - 1 < nrp else 0.0), flush=True)


if __name__ == "__main__":
    main()
