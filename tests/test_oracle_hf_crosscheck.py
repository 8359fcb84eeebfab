"""Second opinion on the timm restatement: transformers' Dinov2WithRegistersModel / SiglipVisionModel are an
independent implementation of the same published architectures.  Reduced depth (3 blocks) at the real widths keeps
this in CPU-seconds; the per-block code path is what is being pinned."""

import pytest
import torch

from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_vit_state_dict
from oracle import vit_oracle

transformers = pytest.importorskip("transformers")
DEPTH = 3


def _pixels(seed=0, batch=1):
    return torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


def test_dinov2_reg4_against_hf():
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    cfg = DINOV2_L14_REG4.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=11, init="stress")
    hf_cfg = Dinov2WithRegistersConfig(hidden_size=cfg.dim, num_hidden_layers=DEPTH, num_attention_heads=cfg.heads,
                                       mlp_ratio=4, image_size=224, patch_size=14, num_register_tokens=4,
                                       layer_norm_eps=1e-6, qkv_bias=True, hidden_act="gelu", layerscale_value=1.0,
                                       attn_implementation="eager")
    m = Dinov2WithRegistersModel(hf_cfg).eval()
    D = cfg.dim
    h = {}
    h["embeddings.patch_embeddings.projection.weight"] = sd["patch_embed.proj.weight"]
    h["embeddings.patch_embeddings.projection.bias"] = sd["patch_embed.proj.bias"]
    h["embeddings.cls_token"] = sd["cls_token"]
    h["embeddings.register_tokens"] = sd["reg_token"]
    h["embeddings.mask_token"] = torch.zeros(1, D)
    # HF adds position_embeddings[:, 0] to cls; timm (no_embed_class) adds nothing → zero that slot
    h["embeddings.position_embeddings"] = torch.cat([torch.zeros(1, 1, D), sd["pos_embed"]], dim=1)
    for i in range(DEPTH):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        wq, wk, wv = sd[p + "attn.qkv.weight"].chunk(3, dim=0)
        bq, bk, bv = sd[p + "attn.qkv.bias"].chunk(3, dim=0)
        for name, w, b in (("query", wq, bq), ("key", wk, bk), ("value", wv, bv)):
            h[q + f"attention.attention.{name}.weight"], h[q + f"attention.attention.{name}.bias"] = w, b
        h[q + "attention.output.dense.weight"], h[q + "attention.output.dense.bias"] = \
            sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        h[q + "norm1.weight"], h[q + "norm1.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        h[q + "norm2.weight"], h[q + "norm2.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        h[q + "layer_scale1.lambda1"], h[q + "layer_scale2.lambda1"] = sd[p + "ls1.gamma"], sd[p + "ls2.gamma"]
        h[q + "mlp.fc1.weight"], h[q + "mlp.fc1.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        h[q + "mlp.fc2.weight"], h[q + "mlp.fc2.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    h["layernorm.weight"], h["layernorm.bias"] = sd["norm.weight"], sd["norm.bias"]
    missing, unexpected = m.load_state_dict(h, strict=False)
    assert not unexpected and not [k for k in missing if "mask_token" not in k], (missing, unexpected)
    x = _pixels()
    with torch.no_grad():
        hs = m(pixel_values=x, output_hidden_states=True).hidden_states
        ref = hs[-2][:, 5:]                       # output of block depth-2, cls + 4 registers dropped
        got = vit_oracle.vit_intermediate(sd, cfg, x)
    assert got.shape == (1, 256, 1024)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_siglip_against_hf():
    from transformers import SiglipVisionConfig, SiglipVisionModel
    cfg = SIGLIP_SO400M_14.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=12, init="stress")
    hf_cfg = SiglipVisionConfig(hidden_size=cfg.dim, intermediate_size=cfg.mlp_hidden, num_hidden_layers=DEPTH,
                                num_attention_heads=cfg.heads, image_size=224, patch_size=14, layer_norm_eps=1e-6,
                                hidden_act="gelu", attn_implementation="eager")
    m = SiglipVisionModel(hf_cfg).eval()
    h = {}
    h["vision_model.embeddings.patch_embedding.weight"] = sd["patch_embed.proj.weight"]
    h["vision_model.embeddings.patch_embedding.bias"] = sd["patch_embed.proj.bias"]
    h["vision_model.embeddings.position_embedding.weight"] = sd["pos_embed"][0]
    for i in range(DEPTH):
        p, q = f"blocks.{i}.", f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[p + "attn.qkv.weight"].chunk(3, dim=0)
        bq, bk, bv = sd[p + "attn.qkv.bias"].chunk(3, dim=0)
        for name, w, b in (("q_proj", wq, bq), ("k_proj", wk, bk), ("v_proj", wv, bv)):
            h[q + f"self_attn.{name}.weight"], h[q + f"self_attn.{name}.bias"] = w, b
        h[q + "self_attn.out_proj.weight"], h[q + "self_attn.out_proj.bias"] = \
            sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        h[q + "layer_norm1.weight"], h[q + "layer_norm1.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        h[q + "layer_norm2.weight"], h[q + "layer_norm2.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        h[q + "mlp.fc1.weight"], h[q + "mlp.fc1.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        h[q + "mlp.fc2.weight"], h[q + "mlp.fc2.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    missing, unexpected = m.load_state_dict(h, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("vision_model.post_layernorm", "vision_model.head")) for k in missing), missing
    x = _pixels(1)
    with torch.no_grad():
        ref = m(pixel_values=x, output_hidden_states=True).hidden_states[-2]
        got = vit_oracle.vit_intermediate(sd, cfg, x)
    assert got.shape == (1, 256, 1152)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_gelu_choice_is_detectable():
    """erf vs tanh GELU must be visible to this cross-check (otherwise it pins nothing)."""
    cfg = SIGLIP_SO400M_14.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=12, init="stress")
    x = _pixels(1)
    a = vit_oracle.vit_intermediate(sd, cfg, x)
    orig = torch.nn.functional.gelu
    try:
        torch.nn.functional.gelu = lambda t: orig(t, approximate="tanh")
        vit_oracle.F.gelu = torch.nn.functional.gelu
        b = vit_oracle.vit_intermediate(sd, cfg, x)
    finally:
        torch.nn.functional.gelu = orig
        vit_oracle.F.gelu = orig
    assert ((a - b).abs().max() / a.abs().max()).item() > 2e-5
