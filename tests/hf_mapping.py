"""hf_mapping.py — build transformers' Dinov2WithRegistersModel / SiglipVisionModel from a timm-keyed state dict
(test infrastructure: the independent second implementation the oracle is cross-checked against)."""

import torch


def build_hf_dinov2_reg4(sd, cfg, depth):
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    hf_cfg = Dinov2WithRegistersConfig(hidden_size=cfg.dim, num_hidden_layers=depth, num_attention_heads=cfg.heads,
                                       mlp_ratio=4, image_size=getattr(cfg, "img_size", 224), patch_size=14, num_register_tokens=4,
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
    for i in range(depth):
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
    return m


def hf_dinov2_penultimate(m, x):
    """output of block depth-2, cls + 4 register tokens dropped (what get_intermediate_layers(n={depth-2}) returns)."""
    with torch.no_grad():
        return m(pixel_values=x, output_hidden_states=True).hidden_states[-2][:, 5:]


def build_hf_siglip(sd, cfg, depth):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    hf_cfg = SiglipVisionConfig(hidden_size=cfg.dim, intermediate_size=cfg.mlp_hidden, num_hidden_layers=depth,
                                num_attention_heads=cfg.heads, image_size=getattr(cfg, "img_size", 224), patch_size=14,
                                layer_norm_eps=1e-6, hidden_act="gelu", attn_implementation="eager")
    m = SiglipVisionModel(hf_cfg).eval()
    h = {}
    h["vision_model.embeddings.patch_embedding.weight"] = sd["patch_embed.proj.weight"]
    h["vision_model.embeddings.patch_embedding.bias"] = sd["patch_embed.proj.bias"]
    h["vision_model.embeddings.position_embedding.weight"] = sd["pos_embed"][0]
    for i in range(depth):
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
    return m


def hf_siglip_penultimate(m, x):
    with torch.no_grad():
        return m(pixel_values=x, output_hidden_states=True).hidden_states[-2]


def build_hf_clip(sd, cfg, depth):
    """OpenAI CLIP ViT-L/14-336 as transformers' CLIPVisionModel: class embedding first, position embedding on every
    token, `pre_layrnorm` (timm norm_pre), conv without bias, quick_gelu."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    hf_cfg = CLIPVisionConfig(hidden_size=cfg.dim, intermediate_size=cfg.mlp_hidden, num_hidden_layers=depth,
                              num_attention_heads=cfg.heads, image_size=cfg.img_size, patch_size=14,
                              layer_norm_eps=1e-6, hidden_act="quick_gelu", attn_implementation="eager")
    m = CLIPVisionModel(hf_cfg).eval()
    h = {}
    h["vision_model.embeddings.class_embedding"] = sd["cls_token"].reshape(-1)
    h["vision_model.embeddings.patch_embedding.weight"] = sd["patch_embed.proj.weight"]
    h["vision_model.embeddings.position_embedding.weight"] = sd["pos_embed"][0]
    h["vision_model.pre_layrnorm.weight"], h["vision_model.pre_layrnorm.bias"] = sd["norm_pre.weight"], sd["norm_pre.bias"]
    for i in range(depth):
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
    assert all(k.startswith("vision_model.post_layernorm") or "position_ids" in k for k in missing), missing
    return m


def hf_clip_penultimate(m, x):
    """output of block depth-2 with the class token dropped."""
    with torch.no_grad():
        return m(pixel_values=x, output_hidden_states=True).hidden_states[-2][:, 1:]
