"""TEST INFRASTRUCTURE ONLY — CPU restatement of multi-scale deformable attention (BASELINE configs[3]).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(aldi_b200/msda.py -> csrc/msda.cu) never does.

What it restates (reference files under
/root/reference/aldi/detr/libs/DeformableDETRDetectron2/deformable_detr/models/ops/):
  * forward  — functions/ms_deform_attn_func.py:41-61 (`ms_deform_attn_core_pytorch`: grid_sample bilinear,
    zero padding, align_corners=False) == src/cuda/ms_deform_im2col_cuda.cuh:34-80 (`..._im2col_bilinear`):
    pixel coordinate = loc * size - 0.5, taps outside the map contribute 0;
  * backward — src/cuda/ms_deform_im2col_cuda.cuh:83-153 (`..._col2im_bilinear`): grad_value scatter of
    w_tap * attn * grad_out, grad_attn = sum_d grad_out * sampled value, grad_loc = size * d(bilinear)/d(coord).

PINNED: tests/golden/make_msda_golden.py imports the reference's own `ms_deform_attn_core_pytorch` (the module's
`import MultiScaleDeformableAttention` is stubbed: that CUDA extension cannot be built here) and stores
its float64 outputs AND autograd gradients for the shapes/seed of the reference's ops/test.py:21-28 plus a
Deformable-DETR-sized case; tests/test_msda_oracle.py replays them against this file.

Layouts (same as the reference op): value (N, S, M, D); spatial_shapes [(H_l, W_l)]; level_start_index [L];
sampling_locations (N, Lq, M, L, P, 2) as (x, y) in [0, 1]; attention_weights (N, Lq, M, L, P);
output (N, Lq, M*D).
"""
import torch


def _taps(loc_l, h, w):
    """loc_l: (N, Lq, M, P, 2).  Returns the four (row, col, weight, valid, d_weight/d_row, d_weight/d_col) taps."""
    x = loc_l[..., 0] * w - 0.5
    y = loc_l[..., 1] * h - 0.5
    inside = (y > -1) & (x > -1) & (y < h) & (x < w)
    y0 = torch.floor(y)
    x0 = torch.floor(x)
    ly, lx = y - y0, x - x0
    hy, hx = 1 - ly, 1 - lx
    y0, x0 = y0.long(), x0.long()
    out = []
    for dy, dx, wt, dwy, dwx in ((0, 0, hy * hx, -hx, -hy), (0, 1, hy * lx, -lx, hy), (1, 0, ly * hx, hx, -ly),
                                 (1, 1, ly * lx, lx, ly)):
        r, c = y0 + dy, x0 + dx
        ok = inside & (r >= 0) & (r <= h - 1) & (c >= 0) & (c <= w - 1)
        out.append((r.clamp(0, h - 1), c.clamp(0, w - 1), wt, ok, dwy, dwx))
    return out


def _gather(value, start, w, r, c):
    """value (N, S, M, D); r, c (N, Lq, M, P) -> (N, Lq, M, P, D)."""
    n, lq, m, p = r.shape
    ni = torch.arange(n).view(n, 1, 1, 1).expand(n, lq, m, p)
    mi = torch.arange(m).view(1, 1, m, 1).expand(n, lq, m, p)
    return value[ni, start + r * w + c, mi]


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    n, s, m, d = value.shape
    lq = sampling_locations.shape[1]
    out = value.new_zeros(n, lq, m, d)
    for l, (h, w) in enumerate(spatial_shapes):
        h, w, start = int(h), int(w), int(level_start_index[l])
        a = attention_weights[:, :, :, l]                                # (N, Lq, M, P)
        for r, c, wt, ok, _, _ in _taps(sampling_locations[:, :, :, l], h, w):
            v = _gather(value, start, w, r, c)                           # (N, Lq, M, P, D)
            out += ((wt * ok * a).unsqueeze(-1) * v).sum(3)
    return out.reshape(n, lq, m * d)


def msda_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output):
    n, s, m, d = value.shape
    lq = sampling_locations.shape[1]
    go = grad_output.reshape(n, lq, m, 1, d)
    g_value = torch.zeros_like(value)
    g_loc = torch.zeros_like(sampling_locations)
    g_attn = torch.zeros_like(attention_weights)
    p = sampling_locations.shape[4]
    ni = torch.arange(n).view(n, 1, 1, 1).expand(n, lq, m, p)
    mi = torch.arange(m).view(1, 1, m, 1).expand(n, lq, m, p)
    for l, (h, w) in enumerate(spatial_shapes):
        h, w, start = int(h), int(w), int(level_start_index[l])
        a = attention_weights[:, :, :, l]
        gy = value.new_zeros(n, lq, m, p)
        gx = value.new_zeros(n, lq, m, p)
        for r, c, wt, ok, dwy, dwx in _taps(sampling_locations[:, :, :, l], h, w):
            v = _gather(value, start, w, r, c)
            dot = (go * v).sum(-1) * ok                                   # sum_d grad_out * tap value
            g_attn[:, :, :, l] += wt * dot
            gy += dwy * dot
            gx += dwx * dot
            contrib = (wt * ok * a).unsqueeze(-1) * go                    # (N, Lq, M, P, D)
            g_value.index_put_((ni, start + r * w + c, mi), contrib, accumulate=True)
        g_loc[:, :, :, l, :, 0] = w * gx * a
        g_loc[:, :, :, l, :, 1] = h * gy * a
    return g_value, g_loc, g_attn
