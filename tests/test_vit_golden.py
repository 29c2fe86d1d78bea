"""`oracle/vit_ref.ViT` -- the ViT trunk `build_vitdet_b_backbone` runs (aldi/backbone.py:21-64) -- pinned to golden vectors
made by an independent published port of detectron2's vit.py: transformers' `VitDetModel` (tests/golden/make_vit_golden.py).
Three cases: zero-padded windows + linearly resampled relative-position tables + bicubic pos_embed on a non-square grid;
the native 14 x 14 grid; several windows with padded edges.  Outputs and every parameter gradient, float64."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_vit_golden as mk  # noqa: E402

from oracle import vit_ref  # noqa: E402

GOLD = torch.load(os.path.join(HERE, "golden", "vit_golden.pt"))


def _oracle(name):
    seed, dim, depth, heads, wblocks, img_size, n, h, w = mk.CASES[name]
    net = vit_ref.ViT(img_size=img_size, embed_dim=dim, depth=depth, num_heads=heads, drop_path_rate=0.0,
                      window_block_indexes=wblocks).double().eval()
    sd, x, gout = mk.vit_case(name)
    net.load_state_dict(sd, strict=True)
    return net, sd, x, gout


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_oracle_vit_matches_the_vitdet_golden(name):
    net, sd, x, gout = _oracle(name)
    y = net(x)["last_feat"]
    ref = GOLD[name]["out"]
    assert y.shape == ref.shape
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-12
    (y * gout).sum().backward()
    named = dict(net.named_parameters())
    for k, sig in GOLD[name]["grads"].items():
        got = mk.grad_signature(k, named[k].grad)
        assert float((got - sig).abs().max()) <= 1e-10 * float(sig[0].abs().clamp_min(1e-30)), (name, k)


def test_live_against_transformers_when_importable():
    """The same comparison without the fixture, when the image carries transformers (it does here and on the GPU boxes)."""
    tr = pytest.importorskip("transformers")
    if not hasattr(tr, "VitDetModel"):
        pytest.skip("this transformers build has no VitDetModel")
    name = "padded_windows_resampled_tables"
    seed, dim, depth, heads, wblocks, img_size, n, h, w = mk.CASES[name]
    cfg = tr.VitDetConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, image_size=img_size,
                          pretrain_image_size=224, patch_size=16, window_block_indices=list(wblocks), window_size=14,
                          use_relative_position_embeddings=True, use_absolute_position_embeddings=True, residual_block_indices=[],
                          drop_path_rate=0.0, mlp_ratio=4, qkv_bias=True, layer_norm_eps=1e-6, hidden_act="gelu")
    hf = tr.VitDetModel(cfg).double().eval()
    net, sd, x, _ = _oracle(name)
    hf.load_state_dict({mk.to_hf(k): v for k, v in sd.items()}, strict=True)
    with torch.no_grad():
        a, b = net(x)["last_feat"], hf(pixel_values=x).last_hidden_state
    assert float((a - b).abs().max() / b.abs().max()) < 1e-12
