"""ConvNeXt bottom-up on B200: forward and explicit backward over a flat parameter buffer (BASELINE
configs[4]).

Mirror of `aldi.backbone.ConvNeXt` (aldi/backbone.py:229-319; blocks :189-227, LayerNorm :321-346, DropPath :160-187)
with the reference's state_dict keys (`downsample_layers.{i}.{j}`, `stages.{i}.{j}.{dwconv,norm,pwconv1,pwconv2,gamma}`,
`norm{i}`), so timm / reference checkpoints load.  Activations are channels-last (N, H, W, C padded to 64); the two
pointwise layers of a block and the stride-k patchify convolutions (as 1x1 layers over space-to-depth rows) run on the
tcgen05 implicit-GEMM kernels (conv_tc / wgrad_tc; fp32 CUDA-core kernels in parity mode), everything else on
csrc/convnext.cu.  DropPath masks are per-(block, sample) factors in {0, 1/keep_prob} drawn on the host
(`draw_keep_masks`) and applied inside the layer-scale + residual kernel.

The bottom-up is pinned against goldens produced by the reference's own class (tests/golden/make_convnext_golden.py,
tests/test_gpu_convnext.py) and returns the four normalised stage outputs the reference hands to Detectron2's FPN;
`DetectorWeights(bottom_up=...)` puts it under the FPN / RPN / box head of `detector.Detector`
(build_convnext_fpn_backbone, aldi/backbone.py:373-392), `B200TrainStep(StepConfig(backbone="convnext", ...))` trains
it with AdamW and EMA (tests/test_gpu_step_parity.py::test_convnext_*).
"""
import os
from collections import OrderedDict

import torch

from . import lib as _l
from . import ops


def _pad64(c):
    return (c + 63) // 64 * 64


class ConvNeXtLayout:
    """state_dict keys <-> ranges of one flat fp32 buffer (every tensor 16-byte aligned)."""

    def __init__(self, depths, dims, in_chans=3):
        self.depths, self.dims = tuple(depths), tuple(dims)
        e = OrderedDict()

        def add(key, shape):
            e[key] = tuple(shape)

        add("downsample_layers.0.0.weight", (dims[0], in_chans, 4, 4)); add("downsample_layers.0.0.bias", (dims[0],))
        add("downsample_layers.0.1.weight", (dims[0],)); add("downsample_layers.0.1.bias", (dims[0],))
        for i in range(3):
            add("downsample_layers.%d.0.weight" % (i + 1), (dims[i],)); add("downsample_layers.%d.0.bias" % (i + 1), (dims[i],))
            add("downsample_layers.%d.1.weight" % (i + 1), (dims[i + 1], dims[i], 2, 2))
            add("downsample_layers.%d.1.bias" % (i + 1), (dims[i + 1],))
        for i, (dep, d) in enumerate(zip(depths, dims)):
            for j in range(dep):
                p = "stages.%d.%d." % (i, j)
                add(p + "gamma", (d,))
                add(p + "dwconv.weight", (d, 1, 7, 7)); add(p + "dwconv.bias", (d,))
                add(p + "norm.weight", (d,)); add(p + "norm.bias", (d,))
                add(p + "pwconv1.weight", (4 * d, d)); add(p + "pwconv1.bias", (4 * d,))
                add(p + "pwconv2.weight", (d, 4 * d)); add(p + "pwconv2.bias", (d,))
            add("norm%d.weight" % i, (d,)); add("norm%d.bias" % i, (d,))
        self.entries = OrderedDict()
        off = 0
        for k, shape in e.items():
            n = 1
            for s in shape:
                n *= s
            self.entries[k] = (off, n, shape)
            off = (off + n + 3) // 4 * 4
        self.numel = off

    @staticmethod
    def to_internal(t):
        """conv weights OIHW -> OHWI (the GEMM's K order (dy, dx, c)); everything else unchanged."""
        return t.permute(0, 2, 3, 1).reshape(-1) if t.dim() == 4 and t.shape[1] != 1 else t.reshape(-1)

    def pack(self, sd):
        flat = torch.zeros(self.numel, dtype=torch.float32)
        missing = [k for k in self.entries if k not in sd]
        if missing:
            raise KeyError("missing keys in state_dict: %s" % missing[:5])
        for k, (off, n, shape) in self.entries.items():
            t = sd[k].detach().to("cpu", torch.float32)
            assert tuple(t.shape) == shape, (k, tuple(t.shape), shape)
            flat[off:off + n] = self.to_internal(t)
        return flat

    def unpack(self, flat):
        flat = flat.detach().cpu()
        out = OrderedDict()
        for k, (off, n, shape) in self.entries.items():
            t = flat[off:off + n]
            if len(shape) == 4 and shape[1] != 1:
                t = t.reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2)
            out[k] = t.reshape(shape).clone()
        return out


def synthetic_state_dict(depths, dims, seed=0, layer_scale_init_value=1e-6, in_chans=3):
    """The reference's own initialisation (aldi/backbone.py:293-296, 201-212): trunc_normal(std 0.02) conv / linear
    weights, zero biases, unit LayerNorms, gamma = layer_scale_init_value — stands in for the ImageNet checkpoint
    (`models/convnext_large_1k_384_backbone.pkl`, configs/Base-RCNN-ConvNeXt-FPN.yaml:3) that cannot be fetched here."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, (_, _, shape) in ConvNeXtLayout(depths, dims, in_chans).entries.items():
        if k.endswith("gamma"):
            sd[k] = torch.full(shape, float(layer_scale_init_value))
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shape)
        elif len(shape) == 1:                       # LayerNorm weights
            sd[k] = torch.ones(shape)
        else:
            sd[k] = torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=g)
    return sd


class _Linear:
    """One dense layer on the implicit-GEMM kernels: packed forward / data-gradient operands + padded bias."""

    def __init__(self, net, key, cout, cin, dgrad=True):
        self.net, self.key, self.cout, self.cin = net, key, cout, cin
        self.cout_p, self.cin_p = _pad64(cout), _pad64(cin)
        dev, dt = net.flat.device, net.dtype
        self.fwd = torch.zeros(self.cout_p, self.cin_p, device=dev, dtype=dt)
        self.bwd = torch.zeros(self.cin_p, self.cout_p, device=dev, dtype=dt) if dgrad else None
        self.bias = torch.zeros(self.cout_p, device=dev)

    def refresh(self):
        w = self.net.view(self.key + ".weight")
        ops.pack_weight(w, self.fwd, cout=self.cout, taps=1, cin=self.cin, cout_p=self.cout_p, cin_p=self.cin_p)
        if self.bwd is not None:
            ops.pack_weight(w, self.bwd, dgrad=True, cout=self.cout, taps=1, cin=self.cin, cout_p=self.cout_p, cin_p=self.cin_p)
        self.bias[:self.cout].copy_(self.net.view(self.key + ".bias"))

    def forward(self, x):
        n, h, w, _ = x.shape
        out = torch.empty(n, h, w, self.cout_p, device=x.device, dtype=x.dtype)
        ops.conv(x, self.fwd, out, bias=self.bias)
        return out

    def backward(self, x, dy, want_dx=True):
        G = self.net.grad
        db = self.net.view(self.key + ".bias", G)
        # bf16: the bias gradient rides on the weight-gradient pass (column sums of the dy tiles it already stages in
        # shared memory) -- the 4x-wide hidden gradient is not read a second time; ALDI_CONVNEXT_COLSUM=1 keeps the
        # separate aldi_colsum pass (A/B knob), which fp32 parity mode always uses
        fused = dy.dtype == torch.bfloat16 and os.environ.get("ALDI_CONVNEXT_COLSUM") != "1"
        ops.wgrad(x, dy, self.net.view(self.key + ".weight", G), cout_store=self.cout, cin_store=self.cin,
                  dbias=db if fused else None)
        if not fused:
            rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
            flat_dy = dy.view(rows, dy.shape[3])
            for c0 in range(0, self.cout, 2048):       # aldi_colsum handles up to 2048 channels per launch
                ops.call("aldi_colsum", flat_dy[:, c0:], self.net.dtc, 1, rows, 0, dy.shape[3], min(2048, self.cout - c0), 1.0,
                         db[c0:])
        if not want_dx:
            return None
        dx = torch.empty_like(x)
        ops.conv(dy, self.bwd, dx, cout_store=self.cin_p)
        return dx


class ConvNeXtBackbone:
    def __init__(self, state_dict, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), drop_path_rate=0.0,
                 out_features=(0, 1, 2, 3), dtype="bf16", device="cuda:0", pixel_mean=(103.53, 116.28, 123.675),
                 pixel_std=(1.0, 1.0, 1.0)):
        _l.load()   # fail loudly if the CUDA library is missing: there is no PyTorch fallback
        self.depths, self.dims, self.out_features = tuple(depths), tuple(dims), tuple(out_features)
        self.dtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.dtc = _l.BF16 if dtype == "bf16" else _l.F32
        self.device = torch.device(device)
        self.layout = ConvNeXtLayout(depths, dims)
        self.flat = self.layout.pack(state_dict).to(self.device)
        self.grad = torch.zeros_like(self.flat)
        self.mean, self.std = tuple(pixel_mean), tuple(pixel_std)
        self.drop_rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]   # aldi/backbone.py:271
        self.lin = OrderedDict()
        self.lin["downsample_layers.0.0"] = _Linear(self, "downsample_layers.0.0", dims[0], 48, dgrad=False)
        for i in range(3):
            self.lin["downsample_layers.%d.1" % (i + 1)] = _Linear(self, "downsample_layers.%d.1" % (i + 1), dims[i + 1], 4 * dims[i])
        for i, (dep, d) in enumerate(zip(depths, dims)):
            for j in range(dep):
                p = "stages.%d.%d." % (i, j)
                self.lin[p + "pwconv1"] = _Linear(self, p + "pwconv1", 4 * d, d)
                self.lin[p + "pwconv2"] = _Linear(self, p + "pwconv2", d, 4 * d)
        self._table = None
        self.refresh()
        self.saved = None

    # ---- flat views ---------------------------------------------------------------------------------------------
    def view(self, key, buf=None):
        off, n, _ = self.layout.entries[key]
        return (self.flat if buf is None else buf)[off:off + n]

    def refresh(self):
        """Re-derive every GEMM operand (forward, data-gradient, padded bias) from the master weights with ONE launch of
        the batched refresh kernel (csrc/optim.cu) — the same descriptor table mechanism as DetectorWeights.refresh."""
        if self._table is None:
            L = _l.load()
            descs = []

            def desc(kind, **kw):
                d = _l.RefreshDesc()
                d.kind, d.out_dtype, d.eps = kind, self.dtc, 1e-6
                for k, v in kw.items():
                    setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
                descs.append(d)

            for lin in self.lin.values():
                w = self.view(lin.key + ".weight")
                geo = dict(cout=lin.cout, taps=1, cin=lin.cin, cout_p=lin.cout_p, cin_p=lin.cin_p)
                desc(0, w=w, out=lin.fwd, **geo)
                if lin.bwd is not None:
                    desc(1, w=w, out=lin.bwd, **geo)
                desc(3, w=self.view(lin.key + ".bias"), out2=lin.bias, cout=lin.cout)
            arr = (_l.RefreshDesc * len(descs))(*descs)
            starts, tot = [], 0
            for d in descs:
                starts.append(tot)
                tot += int(L.aldi_refresh_blocks(_l.ctypes.byref(d)))
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            st = torch.tensor(starts, dtype=torch.int32).to(self.device)
            self._table = (raw, st, len(descs), tot)
        raw, st, n, tot = self._table
        ops.call("aldi_refresh_operands", raw, st, n, tot)

    def state_dict(self):
        return self.layout.unpack(self.flat)

    def draw_keep_masks(self, n, generator=None):
        """Per block the DropPath factor of every sample (aldi/backbone.py:176-181): Bernoulli(keep_prob) / keep_prob."""
        out = []
        for r in self.drop_rates:
            out.append(None if r <= 0 else (torch.rand(n, generator=generator) < (1 - r)).float() / (1 - r))
        return out

    # ---- pieces ---------------------------------------------------------------------------------------------------
    def _ln(self, x, key, c):
        n, h, w, cp = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(n * h * w, 2, device=x.device)
        ops.call("aldi_layernorm_forward", x, self.view(key + ".weight"), self.view(key + ".bias"), 1e-6, n * h * w, c, cp,
                 self.dtc, y, stats)
        return y, stats

    def _ln_bwd(self, x, stats, dy, key, c, dx=None):
        n, h, w, cp = x.shape
        acc = dx is not None
        if dx is None:
            dx = torch.empty_like(x)
        ops.call("aldi_layernorm_backward", x, self.view(key + ".weight"), stats, dy, n * h * w, c, cp, self.dtc, dx, int(acc),
                 self.view(key + ".weight", self.grad), self.view(key + ".bias", self.grad))
        return dx

    # ---- forward (aldi/backbone.py:300-319) ------------------------------------------------------------------------------
    def forward(self, images_u8, sizes, keep_masks=None, save=True):
        """images_u8 (N, 3, H, W) uint8 on the device, H, W multiples of 32; sizes (N, 2) int32 valid (h, w).
        Returns {stage: (N, H/s, W/s, pad64(dim)) channels-last}; keeps what `backward` needs when save=True."""
        n, _, hp, wp = images_u8.shape
        assert hp % 32 == 0 and wp % 32 == 0
        dev, dt = images_u8.device, self.dtype
        keep_masks = keep_masks or [None] * sum(self.depths)
        S = {"blocks": [], "down": [], "norms": {}} if save else None
        patches = torch.empty(n, hp // 4, wp // 4, 64, device=dev, dtype=dt)
        ops.call("aldi_patchify_image", images_u8, sizes, patches, n, hp, wp, 4, 64, self.dtc, ops.host_floats(self.mean),
                 ops.host_floats(self.std))
        u0 = self.lin["downsample_layers.0.0"].forward(patches)
        x, st = self._ln(u0, "downsample_layers.0.1", self.dims[0])
        if save:
            S["stem"] = (patches, u0, st)
        outs, blk = {}, 0
        for i, (dep, d) in enumerate(zip(self.depths, self.dims)):
            if i > 0:
                dprev = self.dims[i - 1]
                ln, st = self._ln(x, "downsample_layers.%d.0" % i, dprev)
                nn_, h, w, cp = ln.shape
                rows = torch.empty(nn_, h // 2, w // 2, _pad64(4 * dprev), device=dev, dtype=dt)
                ops.call("aldi_space_to_depth", ln, rows, nn_, h // 2, w // 2, 2, dprev, cp, rows.shape[3], self.dtc, 0)
                xn = self.lin["downsample_layers.%d.1" % i].forward(rows)
                if save:
                    S["down"].append((x, st, rows))
                x = xn
            nn_, h, w, cp = x.shape
            for j in range(dep):
                p = "stages.%d.%d." % (i, j)
                dw = torch.empty_like(x)
                ops.call("aldi_dwconv7", x, self.view(p + "dwconv.weight"), self.view(p + "dwconv.bias"), nn_, h, w, d, cp, self.dtc,
                         0, dw, 0)
                ln, st = self._ln(dw, p + "norm", d)
                hid = self.lin[p + "pwconv1"].forward(ln)
                act = torch.empty_like(hid)
                ops.call("aldi_gelu", hid, None, act, hid.numel(), self.dtc)
                u = self.lin[p + "pwconv2"].forward(act)
                keep = keep_masks[blk]
                keep = keep.to(dev, torch.float32) if keep is not None else None
                out = torch.empty_like(x)
                ops.call("aldi_layerscale_forward", u, x, self.view(p + "gamma"), keep, nn_ * h * w, h * w, d, cp, self.dtc, out)
                if save:
                    S["blocks"].append((p, x, dw, st, ln, hid, act, u, keep))
                x = out
                blk += 1
            if i in self.out_features:
                o, st = self._ln(x, "norm%d" % i, d)
                outs[i] = o
                if save:
                    S["norms"][i] = (x, st)
            if save:
                S.setdefault("stage_out_shape", {})[i] = x.shape
        self.saved = S
        return outs

    # ---- backward ---------------------------------------------------------------------------------------------------
    def backward(self, d_outs):
        """d_outs: {stage: gradient of that output, same shape / dtype}.  Accumulates every parameter gradient into
        `self.grad` (flat, same layout as the parameters)."""
        S, G = self.saved, self.grad
        assert S is not None, "forward(save=True) first"
        dx = None
        bi = len(S["blocks"])
        for i in reversed(range(4)):
            d = self.dims[i]
            if i in d_outs:
                xs, st = S["norms"][i]
                dx = self._ln_bwd(xs, st, d_outs[i], "norm%d" % i, d, dx=dx)
            if dx is None:
                dx = torch.zeros(S["stage_out_shape"][i], device=self.device, dtype=self.dtype)
            nn_, h, w, cp = dx.shape
            for j in reversed(range(self.depths[i])):
                bi -= 1
                p, x, dw, st, ln, hid, act, u, keep = S["blocks"][bi]
                du = torch.empty_like(u)
                ops.call("aldi_layerscale_backward", u, dx, self.view(p + "gamma"), keep, nn_ * h * w, h * w, d, cp, self.dtc, du,
                         self.view(p + "gamma", G))
                dact = self.lin[p + "pwconv2"].backward(act, du)
                dhid = torch.empty_like(hid)
                ops.call("aldi_gelu", hid, dact, dhid, hid.numel(), self.dtc)
                dln = self.lin[p + "pwconv1"].backward(ln, dhid)
                ddw = self._ln_bwd(dw, st, dln, p + "norm", d)
                ops.call("aldi_dwconv7_wgrad", x, ddw, nn_, h, w, d, cp, self.dtc, self.view(p + "dwconv.weight", G))
                ops.call("aldi_colsum", ddw, self.dtc, 1, nn_ * h * w, 0, cp, d, 1.0, self.view(p + "dwconv.bias", G))
                # residual: d(input) = dx + dwconv^T(ddw), accumulated in place
                ops.call("aldi_dwconv7", ddw, self.view(p + "dwconv.weight"), None, nn_, h, w, d, cp, self.dtc, 1, dx, 1)
            if i > 0:
                xprev, st, rows = S["down"][i - 1]
                drows = self.lin["downsample_layers.%d.1" % i].backward(rows, dx)
                dprev = self.dims[i - 1]
                dln = torch.zeros_like(xprev)
                ops.call("aldi_space_to_depth", drows, dln, nn_, h, w, 2, dprev, xprev.shape[3], drows.shape[3], self.dtc, 1)
                dx = self._ln_bwd(xprev, st, dln, "downsample_layers.%d.0" % i, dprev)
            else:
                patches, u0, st = S["stem"]
                du0 = self._ln_bwd(u0, st, dx, "downsample_layers.0.1", self.dims[0])
                self.lin["downsample_layers.0.0"].backward(patches, du0, want_dx=False)
        self.saved = None
