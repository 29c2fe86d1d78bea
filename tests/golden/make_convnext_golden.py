"""Golden vectors for the ConvNeXt bottom-up from the REFERENCE's own class (authoring container only).

aldi/backbone.py is executed with stand-ins for the Detectron2 names it imports; `ConvNeXt` / `ConvNextBlock` /
`LayerNorm` / `DropPath` (aldi/backbone.py:160-346) are plain torch code.  The model runs in float64, training mode,
with DropPath masks supplied by `convnext_case` (the module-level `drop_path` is replaced by one that consumes them in
call order), and torch.autograd provides every parameter gradient of loss = sum_i <out_i, G_i>.  Weights, inputs, masks
and G_i are regenerated from seeds, so only outputs and gradients are stored.

    python tests/golden/make_convnext_golden.py     ->  tests/golden/convnext_golden.pt
"""
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/aldi/backbone.py"

# name -> (seed, depths, dims, drop_path_rate, layer_scale, N, H, W)
CASES = {
    "tiny": (5, (1, 1, 2, 1), (32, 64, 96, 128), 0.2, 0.5, 2, 64, 96),
    "wide": (6, (2, 1, 1, 1), (96, 192, 384, 768), 0.1, 1e-2, 1, 32, 64),
}
MEAN, STD = (103.53, 116.28, 123.675), (57.375, 57.12, 58.395)


def convnext_state_dict(seed, depths, dims):
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, s=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float64) * s

    sd = {}
    sd["downsample_layers.0.0.weight"], sd["downsample_layers.0.0.bias"] = rn(dims[0], 3, 4, 4, s=0.15), rn(dims[0], s=0.1)
    sd["downsample_layers.0.1.weight"], sd["downsample_layers.0.1.bias"] = 1 + rn(dims[0], s=0.1), rn(dims[0], s=0.1)
    for i in range(3):
        sd["downsample_layers.%d.0.weight" % (i + 1)], sd["downsample_layers.%d.0.bias" % (i + 1)] = 1 + rn(dims[i], s=0.1), rn(dims[i], s=0.1)
        sd["downsample_layers.%d.1.weight" % (i + 1)] = rn(dims[i + 1], dims[i], 2, 2, s=(1.0 / (4 * dims[i])) ** 0.5)
        sd["downsample_layers.%d.1.bias" % (i + 1)] = rn(dims[i + 1], s=0.1)
    for i, (dep, d) in enumerate(zip(depths, dims)):
        for j in range(dep):
            p = "stages.%d.%d." % (i, j)
            sd[p + "gamma"] = 0.5 + rn(d, s=0.2)
            sd[p + "dwconv.weight"], sd[p + "dwconv.bias"] = rn(d, 1, 7, 7, s=0.1), rn(d, s=0.1)
            sd[p + "norm.weight"], sd[p + "norm.bias"] = 1 + rn(d, s=0.1), rn(d, s=0.1)
            sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"] = rn(4 * d, d, s=(1.0 / d) ** 0.5), rn(4 * d, s=0.1)
            sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"] = rn(d, 4 * d, s=(0.25 / d) ** 0.5), rn(d, s=0.1)
        sd["norm%d.weight" % i], sd["norm%d.bias" % i] = 1 + rn(d, s=0.1), rn(d, s=0.1)
    return sd


def convnext_case(name):
    seed, depths, dims, dpr, ls, n, h, w = CASES[name]
    g = torch.Generator().manual_seed(seed + 1000)
    sd = convnext_state_dict(seed, depths, dims)
    images = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
    rates = [x.item() for x in torch.linspace(0, dpr, sum(depths))]
    keeps = []   # per block: (N,) factors in {0, 1/keep_prob}; rate 0 -> None (nn.Identity in the reference)
    for r in rates:
        if r > 0:
            keeps.append((torch.rand(n, generator=g) < (1 - r)).double() / (1 - r))
        else:
            keeps.append(None)
    gouts = [torch.randn(n, d, h // (4 * 2 ** i), w // (4 * 2 ** i), generator=g, dtype=torch.float64) for i, d in enumerate(dims)]
    return sd, images, keeps, gouts


def grad_summary(key, g):
    r = torch.randn(g.numel(), generator=torch.Generator().manual_seed(sum(key.encode())), dtype=torch.float64)
    return (float(g.double().norm()), float((g.double().flatten() * r).sum()))


def load_reference():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Registry:
        def register(self):
            return lambda f: f

    mod("detectron2", model_zoo=None)
    mod("detectron2.config", instantiate=None)
    mod("detectron2.modeling")
    mod("detectron2.modeling.backbone", Backbone=nn.Module)
    mod("detectron2.modeling.backbone.vit", get_vit_lr_decay_rate=None)
    mod("detectron2.modeling.backbone.build", BACKBONE_REGISTRY=Registry())
    mod("detectron2.modeling.backbone.utils", get_abs_pos=None)
    mod("detectron2.modeling.backbone.fpn", FPN=None, LastLevelMaxPool=None)
    mod("detectron2.layers", ShapeSpec=None)
    ns = {"__name__": "ref_aldi_backbone"}
    exec(compile(open(REF).read(), REF, "exec"), ns)
    return ns


def main():
    ns = load_reference()
    out = {}
    for name in CASES:
        seed, depths, dims, dpr, ls, n, h, w = CASES[name]
        sd, images, keeps, gouts = convnext_case(name)
        model = ns["ConvNeXt"](in_chans=3, depths=list(depths), dims=list(dims), drop_path_rate=dpr, layer_scale_init_value=ls,
                               out_features=[0, 1, 2, 3]).double()
        model.load_state_dict(sd, strict=True)
        model.train()
        queue = [k for k in keeps if k is not None]

        def drop_path(x, drop_prob=0.0, training=False, scale_by_keep=True):
            if drop_prob == 0.0 or not training:
                return x
            k = queue.pop(0)
            return x * k.view(-1, 1, 1, 1)

        ns["drop_path"] = drop_path
        x = (images.double() - torch.tensor(MEAN, dtype=torch.float64).view(1, 3, 1, 1)) / torch.tensor(STD, dtype=torch.float64).view(1, 3, 1, 1)
        outs = model(x)
        assert not queue
        loss = sum((outs[i] * gouts[i]).sum() for i in range(4))
        loss.backward()
        grads = {k: p.grad.detach() for k, p in model.named_parameters()}
        out[name] = {"outs": [outs[i].detach().float() for i in range(4)]}
        if name == "tiny":
            out[name]["grads"] = {k: v.float() for k, v in grads.items()}
        else:   # large parameters: keep the norm and one seeded random projection of every gradient
            out[name]["grad_summary"] = {k: grad_summary(k, v) for k, v in grads.items()}
        print(name, [tuple(o.shape) for o in out[name]["outs"]], len(grads))
    torch.save(out, os.path.join(HERE, "convnext_golden.pt"))


if __name__ == "__main__":
    main()
