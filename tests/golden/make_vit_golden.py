"""Golden vectors for the ViT trunk of the ViTDet backbone from an INDEPENDENT published implementation of detectron2's
`modeling/backbone/vit.py`: HuggingFace transformers' `VitDetModel` (transformers 5.5.0 in the authoring image; its
modeling_vitdet.py is a port of detectron2's ViT: `get_absolute_positions` = `get_abs_pos`, `get_rel_pos`,
`add_decomposed_relative_positions`, `window_partition` / `window_unpartition`, `VitDetLayer` = `Block`).

Detectron2 itself is absent from /root/reference (a git dependency without a pinned commit, not installable offline), so the
reference's own class cannot be executed for this path; this is the next best pin for `oracle/vit_ref.ViT`, which is what
`build_vitdet_b_backbone` instantiates under `checkpointed_vit_forward` (aldi/backbone.py:21-64).  SimpleFeaturePyramid has no
counterpart in transformers and stays restated-only.

The model runs in float64, eval mode (DropPath is the identity), on the weight / input seeds of `vit_case`; torch.autograd
provides every parameter gradient of loss = <last_hidden_state, G>.  Outputs are stored whole; of every parameter gradient
the norm and four seeded random projections (`grad_signature`): a 5-number fingerprint per tensor keeps the fixture small.

    python tests/golden/make_vit_golden.py     ->  tests/golden/vit_golden.pt
"""
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (seed, embed_dim, depth, heads, window blocks, ViT img_size, N, H, W)
CASES = {
    "padded_windows_resampled_tables": (3, 128, 4, 2, (0, 2), 128, 2, 128, 160),   # 8 x 10 tokens under 14 x 14 windows; 15-row tables
    "native_grid": (4, 128, 3, 2, (1,), 224, 1, 224, 224),                         # 14 x 14 tokens: pos_embed and tables as stored
    "multi_window": (5, 64, 2, 1, (0,), 128, 1, 256, 320),                         # 16 x 20 tokens: 2 x 2 windows, edge windows padded
}


def vit_case(name):
    """-> (state_dict with the oracle's / Detectron2's key names, input, output-gradient), all float64, from the case's seed."""
    seed, dim, depth, heads, wblocks, img_size, n, h, w = CASES[name]
    g = torch.Generator().manual_seed(seed)
    hd, hidden = dim // heads, 4 * dim

    def rn(*shape, s=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float64) * s

    sd = {"pos_embed": rn(1, 14 * 14 + 1, dim, s=0.3), "patch_embed.proj.weight": rn(dim, 3, 16, 16, s=0.05),
          "patch_embed.proj.bias": rn(dim, s=0.1)}
    for i in range(depth):
        p = "blocks.%d." % i
        s_ = 14 if i in wblocks else img_size // 16
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".weight"], sd[p + nm + ".bias"] = 1 + rn(dim, s=0.1), rn(dim, s=0.1)
        sd[p + "attn.rel_pos_h"], sd[p + "attn.rel_pos_w"] = rn(2 * s_ - 1, hd, s=0.3), rn(2 * s_ - 1, hd, s=0.3)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = rn(3 * dim, dim, s=dim ** -0.5), rn(3 * dim, s=0.1)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = rn(dim, dim, s=dim ** -0.5), rn(dim, s=0.1)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = rn(hidden, dim, s=dim ** -0.5), rn(hidden, s=0.1)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = rn(dim, hidden, s=hidden ** -0.5), rn(dim, s=0.1)
    x = rn(n, 3, h, w)
    gout = rn(n, dim, h // 16, w // 16)
    return sd, x, gout


def grad_signature(key, grad):
    """(norm, 4 projections onto N(0, 1) vectors drawn from a seed derived from the key) of a gradient tensor."""
    g = torch.Generator().manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(key)) % (2 ** 31))
    flat = grad.detach().double().reshape(-1)
    proj = torch.randn(4, flat.numel(), generator=g, dtype=torch.float64)
    return torch.cat([flat.norm().reshape(1), proj @ flat])


def to_hf(key):
    key = key.replace("pos_embed", "embeddings.position_embeddings").replace("patch_embed.proj", "embeddings.projection")
    return key.replace("blocks.", "encoder.layer.").replace(".attn.", ".attention.")


def main():
    import transformers
    from transformers import VitDetConfig, VitDetModel
    out = {"transformers": transformers.__version__}
    for name, (seed, dim, depth, heads, wblocks, img_size, n, h, w) in CASES.items():
        cfg = VitDetConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, image_size=img_size,
                           pretrain_image_size=224, patch_size=16, window_block_indices=list(wblocks), window_size=14,
                           use_relative_position_embeddings=True, use_absolute_position_embeddings=True,
                           residual_block_indices=[], drop_path_rate=0.0, mlp_ratio=4, qkv_bias=True, layer_norm_eps=1e-6,
                           hidden_act="gelu")
        model = VitDetModel(cfg).double().eval()
        sd, x, gout = vit_case(name)
        model.load_state_dict({to_hf(k): v for k, v in sd.items()}, strict=True)
        y = model(pixel_values=x).last_hidden_state
        (y * gout).sum().backward()
        grads = {}
        hf_named = dict(model.named_parameters())
        for k in sd:
            grads[k] = grad_signature(k, hf_named[to_hf(k)].grad)
        out[name] = {"out": y.detach().clone(), "grads": grads}
    torch.save(out, os.path.join(HERE, "vit_golden.pt"))
    print("wrote vit_golden.pt:", {k: tuple(v["out"].shape) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
