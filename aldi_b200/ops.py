"""Thin Python wrappers: torch tensors (device memory + strides) -> C-ABI calls on the current CUDA stream.

torch is used for memory and streams only; all arithmetic happens in libaldi_b200.so.
"""
import ctypes

import torch

from . import lib as _l


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelProfiler:
    """Per-launch device timing with CUDA events on the launching stream (bench.py roofline numbers)."""

    def __init__(self):
        self.records = []  # (name, flops, bytes, key, start_event, end_event)

    def summary(self, by_key=False):
        """Aggregate per entry point (or per (entry point, shape key) with by_key=True)."""
        torch.cuda.synchronize()
        out = {}
        for name, flops, nbytes, key, e0, e1 in self.records:
            k = (name, key) if by_key else name
            d = out.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
            d["launches"] += 1
        return out


_PROFILER = None
_SHADOW = None


def set_shadow(fn):
    """Test seam: fn(kind, run, **call) wraps every tensor-core conv / wgrad call (`kind` = "conv" | "wgrad", `run()`
    performs the real launch, `call` holds the tensors and keyword arguments) so a test can re-run the SAME call on the
    fp32 CUDA-core kernels and compare (tests/test_gpu_fullsize.py shadows every launch of a full-size step)."""
    global _SHADOW
    _SHADOW = fn


def set_profiler(p):
    global _PROFILER
    _PROFILER = p


class KeyRecorder:
    """Records (entry point, flops, bytes, shape key) in launch order without touching the stream (used to label
    the kernels of a CUPTI trace, which follow the same order)."""

    no_events = True

    def __init__(self):
        self.records = []


def _launch(name, fn, flops=0.0, nbytes=0.0, key=None):
    if _PROFILER is None:
        return fn()
    if getattr(_PROFILER, "no_events", False):
        _PROFILER.records.append((name, flops, nbytes, key() if callable(key) else key))
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    _PROFILER.records.append((name, flops, nbytes, key() if callable(key) else key, e0, e1))
    return rc


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dt(t):
    if t.dtype == torch.float32:
        return _l.F32
    if t.dtype == torch.bfloat16:
        return _l.BF16
    raise TypeError("unsupported dtype %s" % t.dtype)


def _cl4(t, what):
    """(N, H, W, C) view with unit channel stride -> (n, h, w, c, sn, sh, sw)."""
    if t.dim() != 4 or t.stride(3) != 1:
        raise ValueError("%s must be a 4-d channels-last view with unit channel stride, got shape %s stride %s"
                         % (what, tuple(t.shape), tuple(t.stride())))
    n, h, w, c = t.shape
    return n, h, w, c, t.stride(0), t.stride(1), t.stride(2)


# ---------------------------------------------------------------------------------------------------
def ema_update(teacher_flat, student_flat, alpha):
    """aldi/ema.py:32-57 on flat fp32 buffers."""
    assert teacher_flat.dtype == torch.float32 and student_flat.dtype == torch.float32
    assert teacher_flat.is_cuda and student_flat.is_cuda and teacher_flat.numel() == student_flat.numel()
    assert teacher_flat.is_contiguous() and student_flat.is_contiguous()
    L = _l.load()
    _l.check(L.aldi_ema_update(_ptr(teacher_flat), _ptr(student_flat), teacher_flat.numel(), float(alpha), _stream()),
             "aldi_ema_update")


def sgd_momentum_step(params, momentum_buf, grads, lr, weight_decay, momentum, grad_scale=1.0, teacher=None,
                      ema_alpha=0.0):
    for t in (params, momentum_buf, grads):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    L = _l.load()
    _l.check(L.aldi_sgd_momentum_step(_ptr(params), _ptr(momentum_buf), _ptr(grads), params.numel(), float(lr),
                                      float(weight_decay), float(momentum), float(grad_scale), _ptr(teacher),
                                      float(ema_alpha), _stream()), "aldi_sgd_momentum_step")


def pack_weight(w, out, *, dgrad=False, scale=None, cout, taps, cin, cout_p, cin_p):
    """fp32 master weight (cout, taps, cin) -> padded GEMM operand in `out` (bf16 or fp32)."""
    assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == cout * taps * cin
    assert out.is_contiguous() and out.numel() == cout_p * taps * cin_p
    L = _l.load()
    _l.check(_launch("aldi_pack_weight", lambda: L.aldi_pack_weight(_ptr(w), _ptr(scale), _ptr(out), _dt(out),
                                                                     int(bool(dgrad)), cout, taps, cin, cout_p, cin_p,
                                                                     _stream())), "aldi_pack_weight")


class Split:
    """An fp32 tensor written as a sum of 2 or 3 bf16 tensors (aldi_split_bf16): `parts[k]` share the shape (and, for
    views made by `as_strided`, the strides) of the tensor they stand for.  The split-bf16 parity mode ("bf16x3",
    "bf16x6") feeds these to the SAME tcgen05 kernels the bf16 step runs."""

    def __init__(self, parts):
        self.parts = list(parts)

    def as_strided(self, size, stride, offset=0):
        """`offset` is relative to the start of each part (torch's as_strided takes an absolute storage offset)."""
        return Split([p.as_strided(size, stride, p.storage_offset() + offset) for p in self.parts])


def split_bf16(x, parts):
    """fp32 channels-last view (1-4 dims, unit last stride) -> Split of contiguous bf16 tensors of the same shape."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and 1 <= x.dim() <= 4
    shape = tuple(x.shape)
    lead = (1,) * (4 - x.dim()) + shape[:-1]
    strides = (0,) * (4 - x.dim()) + tuple(x.stride()[:-1])
    out = torch.empty((parts,) + shape, device=x.device, dtype=torch.bfloat16)
    L = _l.load()
    _l.check(_launch("aldi_split_bf16", lambda: L.aldi_split_bf16(
        _ptr(x), lead[0], lead[1], lead[2], shape[-1], strides[0], strides[1], strides[2], _ptr(out), out[0].numel(), parts,
        _stream())), "aldi_split_bf16")
    return Split([out[k] for k in range(parts)])


def split_view(v, parts):
    """Split that keeps the strides of a non-contiguous view (stride-2 views of 1x1 convs, level slabs of the RPN
    gradient map): the whole underlying allocation is split and the view re-applied to every part, so the kernels
    build the SAME strided TMA tensor maps as in the bf16 step."""
    if v.is_contiguous():
        return split_bf16(v, parts)
    st = v.untyped_storage()
    flat = torch.empty(0, dtype=torch.float32, device=v.device).set_(st, 0, (st.nbytes() // 4,), (1,))
    s = split_bf16(flat, parts)
    return s.as_strided(tuple(v.shape), tuple(v.stride()), v.storage_offset())


def _split_terms(parts):
    """Product terms kept by the split mode, most significant first: (0,0), (1,0), (0,1) for two parts (error
    ~2^-16 per product); three parts add (1,1), (2,0), (0,2) (fp32-level)."""
    return [(i, j) for t in range(parts) for i in range(t + 1) for j in [t - i]]


def _conv_split(x, wp, out, *, taps_h, taps_w, pad_h, pad_w, stride, scale, bias, residual, res_mode, mask, relu,
                accumulate, cout_store):
    """aldi_conv_tc once per product term of (split activations) x (split weights) into ONE fp32 accumulation, then the
    layer's epilogue (aldi_conv_epilogue_f32): the tensor-core main loops of the bf16 step at fp32-level accuracy."""
    assert stride == 1, "tensor-core path: pass a strided view instead of a stride"
    parts = len(wp.parts)
    xs = x if isinstance(x, Split) else split_view(x, parts)
    x0, w0 = xs.parts[0], wp.parts[0]
    n, xh, xw, xc, _, _, _ = _cl4(x0, "x")
    on, oh, ow, oc, o_sn, o_sh, o_sw = _cl4(out, "out")
    assert on == n and out.dtype == torch.float32
    cout_p = w0.shape[0]
    assert w0.numel() == cout_p * taps_h * taps_w * xc, (w0.shape, taps_h, taps_w, xc)
    raw = torch.empty(n, oh, ow, cout_p, device=out.device, dtype=torch.float32)
    L = _l.load()

    def params(xa, wa, o, acc):
        p = _l.ConvParams()
        if xa is not None:
            _, _, _, _, a_sn, a_sh, a_sw = _cl4(xa, "x")
            p.x = xa.data_ptr(); p.x_sw = a_sw; p.x_sh = a_sh; p.x_sn = a_sn
            p.w = wa.data_ptr()
        p.x_c = xc; p.x_w = xw; p.x_h = xh; p.x_n = n
        p.cout_p = cout_p
        p.taps_h = taps_h; p.taps_w = taps_w; p.pad_h = pad_h; p.pad_w = pad_w; p.stride = 1
        p.n = n; p.ho = oh; p.wo = ow
        p.out = o.data_ptr(); p.out_dtype = _l.F32
        _, _, _, _, q_sn, q_sh, q_sw = _cl4(o, "out")
        p.out_sw = q_sw; p.out_sh = q_sh; p.out_sn = q_sn
        p.accumulate = int(acc)
        return p

    for t, (i, j) in enumerate(_split_terms(parts)):
        p = params(xs.parts[i], wp.parts[j], raw, t > 0)
        p.cout_store = cout_p
        _l.check(_launch("aldi_conv_tc", lambda: L.aldi_conv_tc(ctypes.byref(p), _stream())), "aldi_conv_tc(split)")
    p = params(None, None, out, accumulate)
    p.cout_store = int(cout_store if cout_store is not None else min(oc, cout_p))
    assert p.cout_store <= oc
    p.scale = scale.data_ptr() if scale is not None else None
    p.bias = bias.data_ptr() if bias is not None else None
    p.relu = int(bool(relu))
    if residual is not None:
        assert res_mode in (1, 2) and residual.dtype == torch.float32
        _, _, _, _, r_sn, r_sh, r_sw = _cl4(residual, "residual")
        p.residual = residual.data_ptr(); p.res_mode = res_mode
        p.res_sw = r_sw; p.res_sh = r_sh; p.res_sn = r_sn
    if mask is not None:
        assert mask.dtype == torch.float32
        _, mh, mw, _, m_sn, m_sh, m_sw = _cl4(mask, "mask")
        assert (mh, mw) == (oh, ow)
        p.mask = mask.data_ptr(); p.mask_sw = m_sw; p.mask_sh = m_sh; p.mask_sn = m_sn
    _l.check(_launch("aldi_conv_epilogue_f32", lambda: L.aldi_conv_epilogue_f32(_ptr(raw), ctypes.byref(p), _stream())),
             "aldi_conv_epilogue_f32")


def conv(x, wp, out, *, taps_h=1, taps_w=1, pad_h=0, pad_w=0, stride=1, scale=None, bias=None, residual=None,
         res_mode=0, mask=None, relu=False, accumulate=False, cout_store=None, algo_cin=None):
    """Implicit-GEMM conv/linear (forward or data-gradient).  x, out, residual, mask are channels-last
    (N,H,W,C) views; wp is the packed (cout_p, taps*C) operand.  bf16 inputs -> tcgen05 path, fp32 -> CUDA cores;
    a `Split` weight operand (fp32 activations) -> the tcgen05 path in split-bf16 parity mode."""
    if isinstance(wp, Split):
        return _conv_split(x, wp, out, taps_h=taps_h, taps_w=taps_w, pad_h=pad_h, pad_w=pad_w, stride=stride, scale=scale,
                           bias=bias, residual=residual, res_mode=res_mode, mask=mask, relu=relu, accumulate=accumulate,
                           cout_store=cout_store)
    n, xh, xw, xc, x_sn, x_sh, x_sw = _cl4(x, "x")
    on, oh, ow, oc, o_sn, o_sh, o_sw = _cl4(out, "out")
    assert on == n
    cout_p = wp.shape[0]
    assert wp.is_contiguous() and wp.numel() == cout_p * taps_h * taps_w * xc, (wp.shape, taps_h, taps_w, xc)
    assert wp.dtype == x.dtype
    p = _l.ConvParams()
    p.x = x.data_ptr(); p.x_c = xc; p.x_w = xw; p.x_h = xh; p.x_n = n
    p.x_sw = x_sw; p.x_sh = x_sh; p.x_sn = x_sn
    p.w = wp.data_ptr(); p.cout_p = cout_p
    p.taps_h = taps_h; p.taps_w = taps_w; p.pad_h = pad_h; p.pad_w = pad_w; p.stride = stride
    p.n = n; p.ho = oh; p.wo = ow
    p.scale = scale.data_ptr() if scale is not None else None
    p.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        assert res_mode in (1, 2) and residual.dtype == x.dtype
        rn, rh, rw, rc, r_sn, r_sh, r_sw = _cl4(residual, "residual")
        p.residual = residual.data_ptr(); p.res_mode = res_mode
        p.res_sw = r_sw; p.res_sh = r_sh; p.res_sn = r_sn
    else:
        p.residual = None; p.res_mode = 0
    if mask is not None:
        assert mask.dtype == x.dtype
        mn, mh, mw, mc, m_sn, m_sh, m_sw = _cl4(mask, "mask")
        assert (mh, mw) == (oh, ow)
        p.mask = mask.data_ptr(); p.mask_sw = m_sw; p.mask_sh = m_sh; p.mask_sn = m_sn
    else:
        p.mask = None
    p.out = out.data_ptr(); p.out_dtype = _dt(out)
    p.cout_store = int(cout_store if cout_store is not None else min(oc, cout_p))
    assert p.cout_store <= oc
    p.out_sw = o_sw; p.out_sh = o_sh; p.out_sn = o_sn
    p.relu = int(bool(relu)); p.accumulate = int(bool(accumulate))
    L = _l.load()
    flops = 2.0 * n * oh * ow * p.cout_store * taps_h * taps_w * (algo_cin or xc)
    # algorithmic bytes: every operand element touched once (input pixels the taps reach ~ the view itself)
    pix = n * oh * ow
    nbytes = (n * xh * xw * xc + wp.numel()) * x.element_size() + pix * p.cout_store * out.element_size() * \
        (2 if accumulate else 1) + (pix * p.cout_store * 2 // (4 if res_mode == 2 else 1) if residual is not None else 0) + \
        (pix * p.cout_store * 2 if mask is not None else 0)
    key = lambda: "%dx%d c%d->%d px%dx%dx%d%s%s%s%s" % (  # noqa: E731
        taps_h, taps_w, xc, p.cout_store, n, oh, ow, " s2" if x_sw != xc else "", " +res%d" % res_mode if res_mode else "",
        " mask" if mask is not None else "", " acc" if accumulate else "")
    if x.dtype == torch.bfloat16:
        def run():
            _l.check(_launch("aldi_conv_tc", lambda: L.aldi_conv_tc(ctypes.byref(p), _stream()), flops, nbytes, key),
                     "aldi_conv_tc")
        if _SHADOW is not None:
            return _SHADOW("conv", run, x=x, wp=wp, out=out, key=key(), kw=dict(
                taps_h=taps_h, taps_w=taps_w, pad_h=pad_h, pad_w=pad_w, stride=stride, scale=scale, bias=bias,
                residual=residual, res_mode=res_mode, mask=mask, relu=relu, accumulate=accumulate, cout_store=p.cout_store))
        run()
    else:
        _l.check(_launch("aldi_conv_f32", lambda: L.aldi_conv_f32(ctypes.byref(p), _stream()), flops, nbytes, key),
                 "aldi_conv_f32")


def wgrad(x, dy, dw, *, taps_h=1, taps_w=1, pad_h=0, pad_w=0, stride=1, scale=None, cout_store=None,
          cin_store=None, dbias=None, split_parts=0):
    """dw[co, tap, ci] += scale[co] * sum_pixels dy[pixel, co] * x[pixel shifted by tap, ci]  (fp32 atomics).
    split_parts = 2 | 3 with fp32 operands: aldi_wgrad_tc once per product term of the split operands (the result is
    accumulated atomically anyway) -- the split-bf16 parity mode of the tensor-core path."""
    if split_parts:
        assert x.dtype == torch.float32 and dy.dtype == torch.float32 and dbias is None
        xs, ds = split_view(x, split_parts), split_view(dy, split_parts)
        for i, j in _split_terms(split_parts):
            wgrad(xs.parts[i], ds.parts[j], dw, taps_h=taps_h, taps_w=taps_w, pad_h=pad_h, pad_w=pad_w, stride=stride,
                  scale=scale, cout_store=cout_store, cin_store=cin_store)
        return
    n, xh, xw, xc, x_sn, x_sh, x_sw = _cl4(x, "x")
    dn, oh, ow, dc, d_sn, d_sh, d_sw = _cl4(dy, "dy")
    assert dn == n and dy.dtype == x.dtype and dw.dtype == torch.float32 and dw.is_contiguous()
    p = _l.WgradParams()
    p.x = x.data_ptr(); p.x_c = xc; p.x_w = xw; p.x_h = xh; p.x_n = n
    p.x_sw = x_sw; p.x_sh = x_sh; p.x_sn = x_sn
    p.dy = dy.data_ptr(); p.dy_c = dc; p.dy_sw = d_sw; p.dy_sh = d_sh; p.dy_sn = d_sn
    p.n = n; p.ho = oh; p.wo = ow
    p.taps_h = taps_h; p.taps_w = taps_w; p.pad_h = pad_h; p.pad_w = pad_w; p.stride = stride
    p.scale = scale.data_ptr() if scale is not None else None
    p.dw = dw.data_ptr()
    p.dbias = dbias.data_ptr() if (dbias is not None and x.dtype == torch.bfloat16) else None
    p.cout_store = int(cout_store if cout_store is not None else dc)
    p.cin_store = int(cin_store if cin_store is not None else xc)
    assert dw.numel() == p.cout_store * taps_h * taps_w * p.cin_store
    L = _l.load()
    flops = 2.0 * n * oh * ow * p.cout_store * taps_h * taps_w * p.cin_store
    nbytes = (n * xh * xw * xc + n * oh * ow * dc) * x.element_size() + dw.numel() * 8
    key = lambda: "%dx%d c%d->%d px%dx%dx%d%s" % (taps_h, taps_w, p.cin_store, p.cout_store, n, oh, ow,  # noqa: E731
                                                 " s2" if x_sw != xc else "")
    if x.dtype == torch.bfloat16:
        def run():
            _l.check(_launch("aldi_wgrad_tc", lambda: L.aldi_wgrad_tc(ctypes.byref(p), _stream()), flops, nbytes, key),
                     "aldi_wgrad_tc")
        if _SHADOW is not None:
            return _SHADOW("wgrad", run, x=x, dy=dy, dw=dw, key=key(), kw=dict(
                taps_h=taps_h, taps_w=taps_w, pad_h=pad_h, pad_w=pad_w, stride=stride, scale=scale,
                cout_store=p.cout_store, cin_store=p.cin_store))
        run()
    else:
        _l.check(_launch("aldi_wgrad_f32", lambda: L.aldi_wgrad_f32(ctypes.byref(p), _stream()), flops, nbytes, key),
                 "aldi_wgrad_f32")


# ---------------------------------------------------------------------------------------------------
# generic call helper for the remaining entry points (tensors -> device pointers, current stream appended)
# ---------------------------------------------------------------------------------------------------
def _arg(a):
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    return a


def call(name, *args):
    L = _l.load()
    cargs = [_arg(a) for a in args]
    nbytes = 0.0
    if _PROFILER is not None:
        # algorithmic bytes of an elementwise / gather kernel ~ every distinct tensor argument touched once
        seen = set()
        for a in args:
            if isinstance(a, torch.Tensor) and a.data_ptr() not in seen:
                seen.add(a.data_ptr())
                nbytes += a.numel() * a.element_size()
    _l.check(_launch(name, lambda: getattr(L, name)(*cargs, _stream()), 0.0, nbytes), name)


def host_floats(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def make_rpn_levels(shapes, strides, cell_anchors, ch_stride, scale_clamp, min_box_size=0.0):
    """shapes: [(H,W)] per level; cell_anchors: per level list of A [x0,y0,x1,y1] (python floats / tensors)."""
    lv = _l.RpnLevels()
    lv.num_levels = len(shapes)
    lv.num_anchors = len(cell_anchors[0])
    off = 0
    for i, ((h, w), s) in enumerate(zip(shapes, strides)):
        lv.h[i], lv.w[i], lv.stride[i], lv.loc_off[i] = int(h), int(w), int(s), off
        off += int(h) * int(w)
        for a in range(lv.num_anchors):
            for k in range(4):
                lv.cell[i][a][k] = float(cell_anchors[i][a][k])
    lv.total_locs = off
    lv.ch_stride = ch_stride
    lv.scale_clamp = scale_clamp
    lv.min_box_size = min_box_size
    return lv


def roi_align(feats, rois, roi_batch, out=None, dout=None, dfeats=None, *, scales, pooled=7, min_level=2,
              canonical_box_size=224.0, canonical_level=4, num_valid=None):
    """feats: list of (N,H,W,C) contiguous channels-last maps.  Forward if dout is None else backward."""
    p = _l.RoiAlignParams()
    for i, f in enumerate(feats):
        assert f.is_contiguous()
        p.feat[i] = f.data_ptr()
        p.feat_h[i], p.feat_w[i] = f.shape[1], f.shape[2]
        p.scale[i] = float(scales[i])
        if dfeats is not None:
            assert dfeats[i].dtype == torch.float32 and dfeats[i].is_contiguous()
            p.dfeat[i] = dfeats[i].data_ptr()
    p.num_levels = len(feats)
    p.min_level = min_level
    p.canonical_box_size = canonical_box_size
    p.canonical_level = canonical_level
    p.channels = feats[0].shape[3]
    p.pooled = pooled
    p.dtype = _dt(feats[0])
    p.rois = rois.data_ptr()
    p.roi_batch = roi_batch.data_ptr()
    p.num_valid = num_valid.data_ptr() if num_valid is not None else None
    p.num_rois = rois.shape[0]
    L = _l.load()
    if dout is None:
        p.out = out.data_ptr()
        _l.check(_launch("aldi_roi_align_forward", lambda: L.aldi_roi_align_forward(ctypes.byref(p), _stream())),
                 "aldi_roi_align_forward")
    else:
        p.dout = dout.data_ptr()
        _l.check(_launch("aldi_roi_align_backward", lambda: L.aldi_roi_align_backward(ctypes.byref(p), _stream())),
                 "aldi_roi_align_backward")
