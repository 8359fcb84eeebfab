"""vit_oracle.py — fp32 CPU restatement of the reference's featurize+project path.  TEST INFRASTRUCTURE ONLY.

"parity unpinned" at the timm boundary: the ViT arithmetic is timm==0.9.10 (pyproject.toml:45), which is not in
/root/reference and not installable offline; this file restates its published algorithm from the call sites
  prismatic/models/backbones/vision/dinosiglip_vit.py:50-68   timm.create_model(..., num_classes=0, img_size=224)
                                                              .forward = get_intermediate_layers(n={depth-2})
  prismatic/models/backbones/vision/base_vision.py:27-32      unpack_tuple
  prismatic/models/backbones/vision/dinosiglip_vit.py:142-147 per-tower forward + torch.cat(dim=2)
  prismatic/util/nn_utils.py:37-53                            FusedMLPProjector
and is cross-checked against transformers' independent implementations (tests/test_oracle_hf_crosscheck.py).

timm semantics restated:
  PatchEmbed   Conv2d(3, D, 14, stride 14, bias) → flatten(2).transpose(1, 2)              (row-major patches)
  _pos_embed   DINOv2-reg4 (no_embed_class=True): x + pos_embed[1,256,D], THEN cat([cls, reg×4, x]) → 261 tokens
               SigLIP (class_token=False): x + pos_embed → 256 tokens; pos_drop/patch_drop/norm_pre = identity
  Block        x = x + ls1(attn(norm1(x)));  x = x + ls2(mlp(norm2(x)))        LayerNorm eps 1e-6
  Attention    qkv Linear → [B,N,3,H,hd] → permute(2,0,3,1,4) → SDPA(scale hd^-0.5) → transpose → proj Linear
  Mlp          fc1 → nn.GELU() (exact erf) → fc2;   LayerScale x*gamma (DINOv2 only)
  output       block index depth-2, prefix tokens dropped, final norm NOT applied
Round 2 (the other fused backbones of materialize.py:48-49):
  img_size     384 px checkpoints: the strided conv yields floor(384/14) = 27 x 27 patches, the last 6 pixel rows /
               columns are never read; pos_embed has one row per patch
  OpenAI CLIP  (clip_vit.py:15-27, `override_act_layer="quick_gelu"`): pre_norm=True → conv WITHOUT bias and a
               `norm_pre` LayerNorm over all tokens after the embedding; no_embed_class=False → cls is prepended FIRST
               and pos_embed[1, 1+N, D] is added to every token; quick_gelu(x) = x * sigmoid(1.702 x); no LayerScale.
               Cross-checked against transformers' CLIPVisionModel (tests/test_oracle_hf_crosscheck.py).
"""

from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from bridgelang_b200.config import LN_EPS, PATCH, VitConfig


def _block(sd: Dict[str, torch.Tensor], p: str, cfg: VitConfig, x: torch.Tensor) -> torch.Tensor:
    B, N, D = x.shape
    H, hd = cfg.heads, cfg.head_dim
    # attention branch (timm Attention.forward)
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    q, k, v = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    h = (attn @ v).transpose(1, 2).reshape(B, N, D)
    h = F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    if cfg.layer_scale:
        h = h * sd[p + "ls1.gamma"]
    x = x + h
    # MLP branch (timm Mlp.forward)
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = h * torch.sigmoid(1.702 * h) if cfg.act == "quick_gelu" else F.gelu(h)
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if cfg.layer_scale:
        h = h * sd[p + "ls2.gamma"]
    return x + h


def vit_embed(sd: Dict[str, torch.Tensor], cfg: VitConfig, pixels: torch.Tensor) -> torch.Tensor:
    """timm patch_embed + _pos_embed → [B, tokens, D]."""
    x = F.conv2d(pixels, sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"), stride=PATCH)
    x = x.flatten(2).transpose(1, 2)
    prefix = []
    if cfg.class_token:
        prefix.append(sd["cls_token"].expand(x.shape[0], -1, -1))
    if cfg.reg_tokens:
        prefix.append(sd["reg_token"].expand(x.shape[0], -1, -1))
    if cfg.no_embed_class:          # timm _pos_embed: position embedding on the patch tokens only, prefix prepended after
        x = x + sd["pos_embed"]
        x = torch.cat(prefix + [x], dim=1) if prefix else x
    else:                           # prefix first, then pos_embed (which has rows for the prefix tokens too)
        x = torch.cat(prefix + [x], dim=1) if prefix else x
        x = x + sd["pos_embed"]
    if cfg.pre_norm:
        x = F.layer_norm(x, (x.shape[-1],), sd["norm_pre.weight"], sd["norm_pre.bias"], LN_EPS)
    return x


def vit_intermediate(sd: Dict[str, torch.Tensor], cfg: VitConfig, pixels: torch.Tensor,
                     run_all_blocks: bool = False, n_blocks: Optional[int] = None) -> torch.Tensor:
    """get_intermediate_layers(n={depth-2}) → [B, 256, D] patch tokens of block depth-2, no final norm.

    run_all_blocks=True also executes the last block like timm 0.9.10's loop does (its output is discarded);
    that is the "reference as executed" workload the CPU baseline times.  n_blocks overrides how many blocks run
    (per-layer parity checks)."""
    sd = {k: v.to(torch.float32) for k, v in sd.items()}
    x = vit_embed(sd, cfg, pixels.to(torch.float32))
    take = cfg.depth - 2 if n_blocks is None else n_blocks - 1
    out = None
    last = cfg.depth if run_all_blocks else take + 1
    for i in range(last):
        x = _block(sd, f"blocks.{i}.", cfg, x)
        if i == take:
            out = x
    return out[:, cfg.n_prefix:]


def projector_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """nn_utils.py:42-53: Linear → GELU → Linear → GELU → Linear (exact-erf GELU)."""
    h = F.gelu(F.linear(x, sd["projector.0.weight"].float(), sd["projector.0.bias"].float()))
    h = F.gelu(F.linear(h, sd["projector.2.weight"].float(), sd["projector.2.bias"].float()))
    return F.linear(h, sd["projector.4.weight"].float(), sd["projector.4.bias"].float())


def fused_features(dino_sd, dino_cfg: VitConfig, siglip_sd, siglip_cfg: VitConfig,
                   pixel_values: Dict[str, torch.Tensor], run_all_blocks: bool = False) -> torch.Tensor:
    """DinoSigLIPViTBackbone.forward (dinosiglip_vit.py:142-147) → [B, 256, 2176]."""
    d = vit_intermediate(dino_sd, dino_cfg, pixel_values["dino"], run_all_blocks)
    s = vit_intermediate(siglip_sd, siglip_cfg, pixel_values["siglip"], run_all_blocks)
    return torch.cat([d, s], dim=2)


def featurize_project(dino_sd, dino_cfg, siglip_sd, siglip_cfg, proj_sd, pixel_values,
                      run_all_blocks: bool = False) -> torch.Tensor:
    """prismatic.py:367-375: vision_backbone(pixel_values) then projector(...) → [B, 256, llm_dim]."""
    return projector_forward(proj_sd, fused_features(dino_sd, dino_cfg, siglip_sd, siglip_cfg, pixel_values,
                                                     run_all_blocks))
