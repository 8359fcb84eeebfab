"""config.py — static shapes of the two towers and the projector of `prism-dinosiglip-224px`.

Values restate what timm 0.9.10 builds for the ids at prismatic/models/backbones/vision/dinosiglip_vit.py:21-24
(SURVEY.md §8c items 1-3); timm itself is not a dependency of this package.
"""

from __future__ import annotations

from dataclasses import dataclass, replace

IMAGE_SIZE = 224
PATCH = 14
NUM_PATCHES = (IMAGE_SIZE // PATCH) ** 2  # 256
PATCH_K = 3 * PATCH * PATCH               # 588
PATCH_LDK = 592                           # 588 padded so the bf16 row pitch is 16-byte aligned (TMA)
LN_EPS = 1e-6


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)          # timm data cfg of the DINOv2 checkpoints
HALF_MEAN, HALF_STD = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)                               # timm data cfg of the SigLIP checkpoints
OPENAI_CLIP_MEAN, OPENAI_CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


@dataclass(frozen=True)
class VitConfig:
    timm_id: str
    dim: int
    depth: int
    heads: int
    mlp_hidden: int
    class_token: bool
    reg_tokens: int
    layer_scale: bool       # timm init_values is not None
    attn_pool: bool         # SigLIP global_pool='map' head: present in checkpoints, never executed on this path
    img_size: int = IMAGE_SIZE          # timm.create_model(..., img_size=default_image_size): 224 / 336 / 384
    mean: tuple = HALF_MEAN             # timm data config (image transform Normalize)
    std: tuple = HALF_STD
    act: str = "gelu"                   # "gelu" (nn.GELU, erf) | "quick_gelu" (override_act_layer of the OpenAI CLIP towers)
    pre_norm: bool = False              # timm pre_norm=True (CLIP): norm_pre LayerNorm after the embedding, conv without bias
    no_embed_class: bool = True         # False (CLIP): pos_embed has a row for the class token as well

    @property
    def grid(self) -> int:
        """PatchEmbed grid: floor(img_size / 14) — a 384 px frame gives 27 x 27 patches (the last 6 pixel rows/columns
        fall outside every patch, as in timm's strided conv)."""
        return self.img_size // PATCH

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def patch_bias(self) -> bool:
        return not self.pre_norm

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads

    @property
    def n_prefix(self) -> int:
        return int(self.class_token) + self.reg_tokens

    @property
    def tokens(self) -> int:
        return self.num_patches + self.n_prefix

    @property
    def n_needed_blocks(self) -> int:
        """get_intermediate_layers(n={depth-2}) keeps the output of block index depth-2 → run depth-1 blocks."""
        return self.depth - 1

    @property
    def hidden_pad(self) -> int:
        return (self.mlp_hidden + 255) // 256 * 256

    def with_depth(self, depth: int) -> "VitConfig":
        return replace(self, depth=depth)

    def flops_per_image(self) -> int:
        """Algorithmic FLOPs (2·M·N·K) of the needed work, SURVEY.md §8(d)."""
        n, d, h = self.tokens, self.dim, self.mlp_hidden
        block = 6 * n * d * d + 4 * n * n * d + 2 * n * d * d + 4 * n * d * h
        return 2 * self.num_patches * PATCH_K * d + self.n_needed_blocks * block


DINOV2_L14_REG4 = VitConfig("vit_large_patch14_reg4_dinov2.lvd142m", dim=1024, depth=24, heads=16, mlp_hidden=4096,
                            class_token=True, reg_tokens=4, layer_scale=True, attn_pool=False,
                            mean=IMAGENET_MEAN, std=IMAGENET_STD)
SIGLIP_SO400M_14 = VitConfig("vit_so400m_patch14_siglip_224", dim=1152, depth=27, heads=16, mlp_hidden=4304,
                             class_token=False, reg_tokens=0, layer_scale=False, attn_pool=True)

# the other fused backbones of the reference registry (materialize.py:48-49)
#   dinosiglip-vit-so-384px  (dinosiglip_vit.py:26-29): both towers at img_size 384 → 27 x 27 = 729 patches
#   dinoclip-vit-l-336px     (dinoclip_vit.py:22-27):  DINOv2 at 336 + OpenAI CLIP ViT-L/14-336 (quick-GELU, pre-norm)
DINOV2_L14_REG4_384 = replace(DINOV2_L14_REG4, img_size=384)
DINOV2_L14_REG4_336 = replace(DINOV2_L14_REG4, img_size=336)
SIGLIP_SO400M_14_384 = replace(SIGLIP_SO400M_14, timm_id="vit_so400m_patch14_siglip_384", img_size=384)
CLIP_L14_336 = VitConfig("vit_large_patch14_clip_336.openai", dim=1024, depth=24, heads=16, mlp_hidden=4096,
                         class_token=True, reg_tokens=0, layer_scale=False, attn_pool=False, img_size=336,
                         mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, act="quick_gelu", pre_norm=True,
                         no_embed_class=False)

FUSED_DIM = DINOV2_L14_REG4.dim + SIGLIP_SO400M_14.dim  # 2176
LLM_DIM = 4096                                          # Llama-2-7B hidden size


def projector_flops_per_image(fused_dim: int = FUSED_DIM, llm_dim: int = LLM_DIM) -> int:
    return 2 * NUM_PATCHES * (fused_dim * 4 * fused_dim + 4 * fused_dim * llm_dim + llm_dim * llm_dim)


def fused_flops_per_image() -> int:
    """405 208 559 616 for the full-depth prism-dinosiglip-224px path."""
    return DINOV2_L14_REG4.flops_per_image() + SIGLIP_SO400M_14.flops_per_image() + projector_flops_per_image()
