"""weights.py — deterministic random-init state dicts with the reference's key names.

There is no network for checkpoints, so both the benchmark and the parity tests use random weights of the right
architecture:

  * "timm-init":   what timm.create_model(pretrained=False) gives — Linear / pos_embed trunc-normal(σ=0.02), zero
                   biases, LayerNorm 1/0, cls/reg σ=1e-6, LayerScale γ=1e-5, conv default.  Used for throughput.
  * "stress-init": same, but LayerScale γ ~ U(0.05, 1), biases ~ N(0, 0.02), LN weight ~ U(0.5, 1.5), LN bias ~
                   N(0, 0.1).  γ=1e-5 hides block bugs (a broken attention moves the output by 1e-6), so every
                   parity gate uses this one.

Key names follow timm's VisionTransformer as consumed at prismatic/models/backbones/vision/dinosiglip_vit.py:50-58
(`blocks.{i}.ls1.gamma`; the HF twin renames it `scale_factor`, extern/hf/modeling_prismatic.py:52-59) and
prismatic/util/nn_utils.py:42-48 (`projector.{0,2,4}.*`).
"""

from __future__ import annotations

from typing import Dict

import torch

from .config import FUSED_DIM, LLM_DIM, NUM_PATCHES, PATCH, VitConfig


def _tn(gen: torch.Generator, *shape: int, std: float = 0.02) -> torch.Tensor:
    t = torch.empty(*shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


def make_vit_state_dict(cfg: VitConfig, seed: int = 1234, init: str = "stress") -> Dict[str, torch.Tensor]:
    assert init in ("timm", "stress")
    gen = torch.Generator().manual_seed(seed)
    stress = init == "stress"
    D, Hm = cfg.dim, cfg.mlp_hidden
    sd: Dict[str, torch.Tensor] = {}

    def bias(n: int) -> torch.Tensor:
        return torch.randn(n, generator=gen) * 0.02 if stress else torch.zeros(n)

    def ln(prefix: str) -> None:
        sd[prefix + ".weight"] = torch.rand(D, generator=gen) + 0.5 if stress else torch.ones(D)
        sd[prefix + ".bias"] = torch.randn(D, generator=gen) * 0.1 if stress else torch.zeros(D)

    if cfg.class_token:
        sd["cls_token"] = torch.randn(1, 1, D, generator=gen) * (0.02 if stress else 1e-6)
    if cfg.reg_tokens:
        sd["reg_token"] = torch.randn(1, cfg.reg_tokens, D, generator=gen) * (0.02 if stress else 1e-6)
    # no_embed_class=True / class_token=False: one row per patch; CLIP (no_embed_class=False): prefix rows as well
    sd["pos_embed"] = _tn(gen, 1, cfg.num_patches + (0 if cfg.no_embed_class else cfg.n_prefix), D)
    fan_in = 3 * PATCH * PATCH
    bound = 1.0 / fan_in ** 0.5                             # nn.Conv2d default (kaiming_uniform a=sqrt(5))
    sd["patch_embed.proj.weight"] = (torch.rand(D, 3, PATCH, PATCH, generator=gen) * 2 - 1) * bound
    patch_bias = (torch.rand(D, generator=gen) * 2 - 1) * bound
    if cfg.patch_bias:
        sd["patch_embed.proj.bias"] = patch_bias
    if cfg.pre_norm:
        ln("norm_pre")
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        ln(p + "norm1")
        sd[p + "attn.qkv.weight"] = _tn(gen, 3 * D, D)
        sd[p + "attn.qkv.bias"] = bias(3 * D)
        sd[p + "attn.proj.weight"] = _tn(gen, D, D)
        sd[p + "attn.proj.bias"] = bias(D)
        if cfg.layer_scale:
            sd[p + "ls1.gamma"] = torch.rand(D, generator=gen) * 0.95 + 0.05 if stress else torch.full((D,), 1e-5)
        ln(p + "norm2")
        sd[p + "mlp.fc1.weight"] = _tn(gen, Hm, D)
        sd[p + "mlp.fc1.bias"] = bias(Hm)
        sd[p + "mlp.fc2.weight"] = _tn(gen, D, Hm)
        sd[p + "mlp.fc2.bias"] = bias(D)
        if cfg.layer_scale:
            sd[p + "ls2.gamma"] = torch.rand(D, generator=gen) * 0.95 + 0.05 if stress else torch.full((D,), 1e-5)
    ln("norm")                                              # final norm: in every checkpoint, never applied here
    if cfg.attn_pool:                                       # SigLIP MAP head: loaded and ignored
        sd["attn_pool.latent"] = _tn(gen, 1, 1, D)
        for n, shape in (("q", (D, D)), ("kv", (2 * D, D)), ("proj", (D, D))):
            sd[f"attn_pool.{n}.weight"] = _tn(gen, *shape)
            sd[f"attn_pool.{n}.bias"] = torch.zeros(shape[0])
        sd["attn_pool.norm.weight"], sd["attn_pool.norm.bias"] = torch.ones(D), torch.zeros(D)
        sd["attn_pool.mlp.fc1.weight"], sd["attn_pool.mlp.fc1.bias"] = _tn(gen, Hm, D), torch.zeros(Hm)
        sd["attn_pool.mlp.fc2.weight"], sd["attn_pool.mlp.fc2.bias"] = _tn(gen, D, Hm), torch.zeros(D)
    return sd


def make_projector_state_dict(fused_dim: int = FUSED_DIM, llm_dim: int = LLM_DIM, seed: int = 4321,
                              init: str = "stress") -> Dict[str, torch.Tensor]:
    """Keys of FusedMLPProjector.projector = nn.Sequential(Linear, GELU, Linear, GELU, Linear)."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    dims = ((0, fused_dim, 4 * fused_dim), (2, 4 * fused_dim, llm_dim), (4, llm_dim, llm_dim))
    for idx, fin, fout in dims:
        bound = 1.0 / fin ** 0.5                            # nn.Linear default init
        sd[f"projector.{idx}.weight"] = (torch.rand(fout, fin, generator=gen) * 2 - 1) * bound
        sd[f"projector.{idx}.bias"] = (torch.rand(fout, generator=gen) * 2 - 1) * bound
        if init == "stress":
            sd[f"projector.{idx}.bias"] = torch.randn(fout, generator=gen) * 0.02
    return sd


def synthetic_frames(batch: int, seed: int = 0, size: int = 224) -> torch.Tensor:
    """uint8 [B,size,size,3] frames, mirroring vla-scripts/extern/verify_openvla.py:74."""
    gen = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), dtype=torch.uint8, generator=gen)


def normalize_for(cfg: VitConfig, frames_u8: torch.Tensor) -> torch.Tensor:
    """ToTensor + Normalize with a tower's own mean/std: fp32 [B,3,S,S]."""
    x = frames_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    m = torch.tensor(cfg.mean, device=x.device).view(1, 3, 1, 1)
    s = torch.tensor(cfg.std, device=x.device).view(1, 3, 1, 1)
    return ((x - m) / s).contiguous()


DINO_MEAN, DINO_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # timm data cfg of the DINOv2 checkpoint
SIGLIP_MEAN, SIGLIP_STD = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)           # timm data cfg of the SigLIP checkpoint


def normalize_frames(frames_u8: torch.Tensor) -> Dict[str, torch.Tensor]:
    """to_tensor + per-tower Normalize of DinoSigLIPImageTransform (dinosiglip_vit.py:33-40): fp32 [B,3,224,224]."""
    x = frames_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    out = {}
    for name, mean, std in (("dino", DINO_MEAN, DINO_STD), ("siglip", SIGLIP_MEAN, SIGLIP_STD)):
        m = torch.tensor(mean, device=x.device).view(1, 3, 1, 1)
        s = torch.tensor(std, device=x.device).view(1, 3, 1, 1)
        out[name] = ((x - m) / s).contiguous()             # NCHW-contiguous, like torchvision's ToTensor output
    return out
