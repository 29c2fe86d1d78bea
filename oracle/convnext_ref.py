"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ConvNeXt bottom-up (aldi/backbone.py:160-346).

Only tests/ may import this.  PINNED: tests/test_convnext_oracle.py replays tests/golden/convnext_golden.pt, which
tests/golden/make_convnext_golden.py produced by executing the reference's own `ConvNeXt` class (float64 outputs and
autograd gradients).  Same module tree and parameter names as the reference (`downsample_layers`, `stages`, `norm{i}`),
so the reference's state_dicts load with strict=True; exposes the `_out_feature_*` tables Detectron2's FPN reads
(aldi/backbone.py:283-284) so that oracle/d2_rcnn.FPN can sit on top of it (build_convnext_fpn_backbone, :373-392).
DropPath takes its per-sample factors from `keep_queue` (filled by the test) instead of `bernoulli_`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LayerNorm(nn.Module):   # aldi/backbone.py:321-346
    def __init__(self, c, eps=1e-6, data_format="channels_last"):
        super().__init__()
        self.weight, self.bias = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c))
        self.eps, self.data_format, self.c = eps, data_format, c

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, (self.c,), self.weight, self.bias, self.eps)
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class Block(nn.Module):       # aldi/backbone.py:189-227
    def __init__(self, dim, drop_path, layer_scale_init_value, owner):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim))
        self.drop_prob = drop_path
        self._owner = [owner]   # not a submodule

    def forward(self, x):
        inp = x
        x = self.dwconv(x).permute(0, 2, 3, 1)
        x = self.pwconv2(F.gelu(self.pwconv1(self.norm(x))))
        x = (self.gamma * x).permute(0, 3, 1, 2)
        if self.drop_prob > 0.0 and self.training:      # aldi/backbone.py:160-181 drop_path
            keep = self._owner[0].keep_queue.pop(0)
            x = x * keep.to(x.dtype).view(-1, 1, 1, 1)
        return inp + x


class ConvNeXt(nn.Module):    # aldi/backbone.py:229-319
    def __init__(self, in_chans=3, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), drop_path_rate=0.0,
                 layer_scale_init_value=1e-6, out_features=(0, 1, 2, 3)):
        super().__init__()
        self.keep_queue = []
        self.downsample_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4),
                                                              LayerNorm(dims[0], data_format="channels_first"))])
        for i in range(3):
            self.downsample_layers.append(nn.Sequential(LayerNorm(dims[i], data_format="channels_first"),
                                                        nn.Conv2d(dims[i], dims[i + 1], kernel_size=2, stride=2)))
        rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.stages = nn.ModuleList()
        cur = 0
        self._out_features = list(out_features)
        self._out_feature_channels, self._out_feature_strides = {}, {}
        for i in range(4):
            self.stages.append(nn.Sequential(*[Block(dims[i], rates[cur + j], layer_scale_init_value, self) for j in range(depths[i])]))
            cur += depths[i]
            self._out_feature_channels[i] = dims[i]
            self._out_feature_strides[i] = 4 * 2 ** i
        for i in range(4):
            self.add_module("norm%d" % i, LayerNorm(dims[i], data_format="channels_first"))

    def forward(self, x):
        outs = {}
        for i in range(4):
            x = self.stages[i](self.downsample_layers[i](x))
            if i in self._out_features:
                outs[i] = getattr(self, "norm%d" % i)(x).contiguous()
        return outs
