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

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads

    @property
    def n_prefix(self) -> int:
        return int(self.class_token) + self.reg_tokens

    @property
    def tokens(self) -> int:
        return NUM_PATCHES + self.n_prefix

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
        return 2 * NUM_PATCHES * PATCH_K * d + self.n_needed_blocks * block


DINOV2_L14_REG4 = VitConfig("vit_large_patch14_reg4_dinov2.lvd142m", dim=1024, depth=24, heads=16, mlp_hidden=4096,
                            class_token=True, reg_tokens=4, layer_scale=True, attn_pool=False)
SIGLIP_SO400M_14 = VitConfig("vit_so400m_patch14_siglip_224", dim=1152, depth=27, heads=16, mlp_hidden=4304,
                             class_token=False, reg_tokens=0, layer_scale=False, attn_pool=True)

FUSED_DIM = DINOV2_L14_REG4.dim + SIGLIP_SO400M_14.dim  # 2176
LLM_DIM = 4096                                          # Llama-2-7B hidden size


def projector_flops_per_image(fused_dim: int = FUSED_DIM, llm_dim: int = LLM_DIM) -> int:
    return 2 * NUM_PATCHES * (fused_dim * 4 * fused_dim + 4 * fused_dim * llm_dim + llm_dim * llm_dim)


def fused_flops_per_image() -> int:
    """405 208 559 616 for the full-depth prism-dinosiglip-224px path."""
    return DINOV2_L14_REG4.flops_per_image() + SIGLIP_SO400M_14.flops_per_image() + projector_flops_per_image()
