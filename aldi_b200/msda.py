"""Multi-scale deformable attention on B200: host mirror of the reference's operator and module.

`MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index, sampling_locations,
attention_weights, im2col_step)` has the signature, argument meaning, gradients and error behaviour of the
reference's autograd wrapper (aldi/detr/libs/DeformableDETRDetectron2/deformable_detr/models/ops/functions/
ms_deform_attn_func.py:21-38) and `MSDeformAttn` that of its module (modules/ms_deform_attn.py:31-115), but the op
runs the hand-written sm_100a kernels of csrc/msda.cu through the C ABI (include/aldi_b200.h aldi_msda_*).  There
is no PyTorch fallback: a missing library raises `AldiError`, CPU tensors raise.
"""
import ctypes
import math
import warnings

import torch
from torch import nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import lib as _l
from . import ops

_DT = {torch.float32: 0, torch.float64: 2}


def _params(value, shapes, starts, loc, attn):
    if not (value.is_cuda and loc.is_cuda and attn.is_cuda):
        raise _l.AldiError("MSDeformAttn runs on the GPU only (the reference's CPU path is `AT_ERROR(\"Not implement on "
                           "cpu\")`, src/cpu/ms_deform_attn_cpu.cpp:17-27)")
    if value.dtype not in _DT or loc.dtype != value.dtype or attn.dtype != value.dtype:
        raise TypeError("MSDeformAttn: value / sampling_locations / attention_weights must share dtype float32 or "
                        "float64, got %s %s %s" % (value.dtype, loc.dtype, attn.dtype))
    n, s, m, d = value.shape
    n2, lq, m2, l, p, two = loc.shape
    assert (n2, m2, two) == (n, m, 2) and tuple(attn.shape) == (n, lq, m, l, p), "MSDeformAttn: inconsistent shapes"
    sh = [(int(h), int(w)) for h, w in (shapes.tolist() if isinstance(shapes, torch.Tensor) else shapes)]
    st = [int(v) for v in (starts.tolist() if isinstance(starts, torch.Tensor) else starts)]
    assert len(sh) == l and len(st) == l
    q = _l.MsdaParams()
    q._keep = ((ctypes.c_int * l)(*[h for h, _ in sh]), (ctypes.c_int * l)(*[w for _, w in sh]), (ctypes.c_int * l)(*st))
    q.spatial_h, q.spatial_w, q.level_start = q._keep
    q.value, q.sampling_loc, q.attn_weight = value.data_ptr(), loc.data_ptr(), attn.data_ptr()
    q.n, q.s, q.m, q.d, q.lq, q.l, q.p = n, s, m, d, lq, l, p
    q.dtype = _DT[value.dtype]
    return q


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        # the reference batches its launches by im2col_step and asserts divisibility (ms_deform_attn_cuda.cu:49-51)
        n = value.shape[0]
        step = min(n, int(im2col_step))
        if n % step != 0:
            raise AssertionError("batch(%d) must divide im2col_step(%d)" % (n, step))
        value, loc, attn = value.contiguous(), sampling_locations.contiguous(), attention_weights.contiguous()
        q = _params(value, value_spatial_shapes, value_level_start_index, loc, attn)
        out = torch.empty(n, loc.shape[1], value.shape[2] * value.shape[3], device=value.device, dtype=value.dtype)
        q.out = out.data_ptr()
        L = _l.load()
        _l.check(L.aldi_msda_forward(ctypes.byref(q), _stream()), "aldi_msda_forward")
        ctx.save_for_backward(value, loc, attn)
        ctx.geometry = (value_spatial_shapes, value_level_start_index)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, loc, attn = ctx.saved_tensors
        shapes, starts = ctx.geometry
        grad_output = grad_output.contiguous()
        q = _params(value, shapes, starts, loc, attn)
        grad_value = torch.zeros_like(value)
        grad_loc = torch.empty_like(loc)
        grad_attn = torch.empty_like(attn)
        q.grad_out, q.grad_value, q.grad_loc, q.grad_attn = (grad_output.data_ptr(), grad_value.data_ptr(),
                                                             grad_loc.data_ptr(), grad_attn.data_ptr())
        L = _l.load()
        _l.check(L.aldi_msda_backward(ctypes.byref(q), _stream()), "aldi_msda_backward")
        return grad_value, None, None, grad_loc, grad_attn, None


class _LinearFunction(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM kernels (aldi_conv_tc / aldi_wgrad_tc as a 1x1 layer over a 1 x 1 x M x C image)
    with fp32 operands split into `parts` bf16 tensors: 3 = fp32-level (the module is an fp32 module in the reference),
    1 = plain bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, parts):
        cout, cin = weight.shape
        cout_p = (cout + 63) // 64 * 64
        m = x.numel() // cin
        x2 = x.reshape(1, 1, m, cin).contiguous().float()
        wp = torch.empty(cout_p, cin, device=x.device)
        ops.pack_weight(weight.detach().contiguous().float(), wp, cout=cout, taps=1, cin=cin, cout_p=cout_p, cin_p=cin)
        bp = torch.zeros(cout_p, device=x.device)
        bp[:cout] = bias.detach()
        out = torch.empty(1, 1, m, cout, device=x.device)
        ops.conv(x2, ops.split_bf16(wp, parts), out, bias=bp, cout_store=cout)
        ctx.save_for_backward(x2, weight)
        ctx.parts, ctx.shape = parts, x.shape
        return out.view(*x.shape[:-1], cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        x2, weight = ctx.saved_tensors
        parts = ctx.parts
        cout, cin = weight.shape
        cout_p = (cout + 63) // 64 * 64
        m = x2.shape[2]
        g2 = torch.zeros(1, 1, m, cout_p, device=grad.device)
        g2[..., :cout] = grad.reshape(1, 1, m, cout)
        wt = torch.empty(cin, cout_p, device=grad.device)          # [cin][cout_p] = W^T, zero padded
        ops.pack_weight(weight.detach().contiguous().float(), wt, dgrad=True, cout=cout, taps=1, cin=cin, cout_p=cout_p, cin_p=cin)
        dx = torch.empty(1, 1, m, cin, device=grad.device)
        ops.conv(g2, ops.split_bf16(wt, parts), dx, cout_store=cin)
        dw = torch.zeros(cout, cin, device=grad.device)
        ops.wgrad(x2, g2, dw.view(-1), cout_store=cout, cin_store=cin, split_parts=parts)
        db = torch.zeros(cout, device=grad.device)
        ops.call("aldi_colsum", g2, _l.F32, 1, m, 0, cout_p, cout, 1.0, db)
        return dx.view(ctx.shape), dw, db, None


class Linear(nn.Module):
    """nn.Linear's parameters (`weight`, `bias`: the reference's checkpoints load) over `_LinearFunction`."""

    def __init__(self, in_features, out_features, parts=3):
        super().__init__()
        if in_features % 64:
            raise ValueError("aldi_b200.msda.Linear: in_features must be a multiple of 64 (tensor-core K blocks), got %d"
                             % in_features)
        self.in_features, self.out_features, self.parts = in_features, out_features, parts
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        if not x.is_cuda:
            raise _l.AldiError("aldi_b200.msda.Linear runs on the GPU only (there is no CPU fallback)")
        return _LinearFunction.apply(x, self.weight, self.bias, self.parts)


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    """modules/ms_deform_attn.py:31-115.  The four projections run on this library's tcgen05 GEMM kernels (`Linear`
    above: split-bf16, fp32-level; no cuBLAS), the sampling itself is MSDeformAttnFunction.  Parameter names match the
    reference so its checkpoints load."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-2 head dimension keeps every bilinear tap one aligned line")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = Linear(d_model, d_model)
        self.output_proj = Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # modules/ms_deform_attn.py:64-78: offsets start as a ring of directions scaled by the point index
        nn.init.constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        n, len_q, _ = query.shape
        n, len_in, _ = input_flatten.shape
        shapes = input_spatial_shapes
        assert int((shapes[:, 0] * shapes[:, 1]).sum()) == len_in
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(n, len_in, self.n_heads, self.d_model // self.n_heads)
        offsets = self.sampling_offsets(query).view(n, len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(n, len_q, self.n_heads, self.n_levels * self.n_points)
        weights = F.softmax(weights, -1).view(n, len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = reference_points[:, :, None, :, None, :2] + \
                offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        out = MSDeformAttnFunction.apply(value, shapes, input_level_start_index, loc, weights, self.im2col_step)
        return self.output_proj(out)
