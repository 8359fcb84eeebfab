"""bridgelang_b200 — B200-native (sm_100a) visual-prefix hot path of OpenVLA / prismatic (CliffKai/BridgeLang).

Host side: Python mirrors of the reference's `VisionBackbone` / projector / `ActionTokenizer` /
`predict_action` interfaces.  Device side: hand-written CUDA kernels behind the C ABI in
include/bridgelang_b200.h (libbridgelang_b200.so, built in-tree by `python -m bridgelang_b200.build`).
"""

from .action_tokenizer import ActionTokenizer
from .config import DINOV2_L14_REG4, FUSED_DIM, LLM_DIM, SIGLIP_SO400M_14, VitConfig, fused_flops_per_image
from .pipeline import PrefixGatherer, VisualPrefixEncoder, gather_prefixes, shard_bounds, shard_pixel_values
from .projector import FusedMLPProjector, PrismaticProjector
from .vision import (CLIPViTBackbone, DinoCLIPImageTransform, DinoCLIPViTBackbone, DinoSigLIPImageTransform,
                     DinoSigLIPViTBackbone, DinoV2ViTBackbone, PrismaticImageProcessor, PrismaticVisionBackbone,
                     SigLIPViTBackbone, VisionBackbone, VisionTransformer)
from .vla import LLMBackbone, OpenVLA, OpenVLAForActionPrediction, PurePromptBuilder, decode_tail_from_logits

__all__ = [
    "ActionTokenizer", "DINOV2_L14_REG4", "SIGLIP_SO400M_14", "VitConfig", "FUSED_DIM", "LLM_DIM",
    "fused_flops_per_image", "VisualPrefixEncoder", "PrefixGatherer", "gather_prefixes", "shard_bounds", "shard_pixel_values",
    "FusedMLPProjector", "PrismaticProjector", "DinoSigLIPImageTransform", "DinoSigLIPViTBackbone", "DinoCLIPViTBackbone", "DinoCLIPImageTransform", "CLIPViTBackbone",
    "DinoV2ViTBackbone", "SigLIPViTBackbone", "PrismaticImageProcessor", "PrismaticVisionBackbone", "VisionBackbone", "VisionTransformer",
    "LLMBackbone", "OpenVLA", "OpenVLAForActionPrediction", "PurePromptBuilder", "decode_tail_from_logits",
]
