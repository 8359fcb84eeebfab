"""vision.py — B200-native mirrors of the reference's vision-backbone classes for `prism-dinosiglip-224px`.

  VisionBackbone / DinoSigLIPViTBackbone   prismatic/models/backbones/vision/base_vision.py:54-90,
                                           prismatic/models/backbones/vision/dinosiglip_vit.py:43-163
  PrismaticVisionBackbone (HF twin)        prismatic/extern/hf/modeling_prismatic.py:63-123
  SigLIPViTBackbone / DinoV2ViTBackbone    prismatic/models/backbones/vision/{siglip_vit,dinov2_vit}.py

Same constructor arguments, attributes, state-dict key names (timm's) and forward contracts as the reference, so
reference checkpoints load with `load_state_dict` and callers (`PrismaticVLM.forward`, prismatic.py:367-375) do not
change.  The arithmetic is NOT timm/PyTorch: `forward` hands raw device pointers to libbridgelang_b200.so
(hand-written sm_100a kernels).  Inference only (the reference's `freeze_vision_backbone` stages and every
`predict_action` use); there is no eager/CPU fallback — a missing library or a CPU tensor raises.
"""

from __future__ import annotations

import ctypes as C
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from functools import partial
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .config import (DINOV2_L14_REG4, LN_EPS, NUM_PATCHES, PATCH, PATCH_K, PATCH_LDK, SIGLIP_SO400M_14, VitConfig)
from .weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD

# Registry =>> same ids as dinosiglip_vit.py:21-30 / siglip_vit.py:8-13 / dinov2_vit.py:9-10 (224 px members)
DINOSigLIP_VISION_BACKBONES = {
    "dinosiglip-vit-so-224px": {"dino": DINOV2_L14_REG4, "siglip": SIGLIP_SO400M_14},
}
SIGLIP_VISION_BACKBONES = {"siglip-vit-so400m": SIGLIP_SO400M_14}
DINOv2_VISION_BACKBONES = {"dinov2-vit-l": DINOV2_L14_REG4}
TIMM_ID_TO_CONFIG = {c.timm_id: c for c in (DINOV2_L14_REG4, SIGLIP_SO400M_14)}


def unpack_tuple(fn: Callable[[Any], Tuple[Any]]) -> Callable[[Any], Any]:
    """base_vision.py:27-32 (kept for API parity; the native towers return a tensor already)."""
    def wrapper(*args: Any, **kwargs: Any) -> Any:
        result = fn(*args, **kwargs)
        return result[0] if isinstance(result, tuple) else result

    return wrapper


# ======================================================================================================
# Parameter containers with timm's module tree (names only — no timm code runs)
# ======================================================================================================
class LayerScale(nn.Module):
    """timm LayerScale: holds `gamma` (or `scale_factor` in the HF twin, modeling_prismatic.py:52-59)."""

    def __init__(self, dim: int, init_values: float = 1e-5, param_name: str = "gamma") -> None:
        super().__init__()
        self.param_name = param_name
        setattr(self, param_name, nn.Parameter(init_values * torch.ones(dim)))

    @property
    def value(self) -> torch.Tensor:
        return getattr(self, self.param_name)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):  # accept either spelling
        other = "scale_factor" if self.param_name == "gamma" else "gamma"
        if prefix + other in state_dict and prefix + self.param_name not in state_dict:
            state_dict[prefix + self.param_name] = state_dict.pop(prefix + other)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _Attention(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)


class Block(nn.Module):
    """Parameter holder named like timm.models.vision_transformer.Block (also the FSDP wrap unit)."""

    def __init__(self, cfg: VitConfig, ls_param_name: str) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(cfg.dim, eps=LN_EPS)
        self.attn = _Attention(cfg.dim)
        self.ls1 = LayerScale(cfg.dim, param_name=ls_param_name) if cfg.layer_scale else nn.Identity()
        self.norm2 = nn.LayerNorm(cfg.dim, eps=LN_EPS)
        self.mlp = _Mlp(cfg.dim, cfg.mlp_hidden)
        self.ls2 = LayerScale(cfg.dim, param_name=ls_param_name) if cfg.layer_scale else nn.Identity()


class _PatchEmbed(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH, bias=True)


class _AttentionPool(nn.Module):
    """SigLIP's MAP head: present in checkpoints (`attn_pool.*`), never executed on this path (num_classes=0 and
    get_intermediate_layers bypass it).  Kept so that load_state_dict(strict=True) of reference weights works."""

    def __init__(self, cfg: VitConfig) -> None:
        super().__init__()
        D = cfg.dim
        self.latent = nn.Parameter(torch.zeros(1, 1, D))
        self.q = nn.Linear(D, D)
        self.kv = nn.Linear(D, 2 * D)
        self.proj = nn.Linear(D, D)
        self.norm = nn.LayerNorm(D, eps=LN_EPS)
        self.mlp = _Mlp(D, cfg.mlp_hidden)


class VisionTransformer(nn.Module):
    """One tower.  `forward(pixels[B,3,224,224]) -> [B,256,D]` == timm `get_intermediate_layers(n={depth-2})`
    after `unpack_tuple` (second-to-last block, prefix tokens dropped, final norm not applied)."""

    def __init__(self, cfg: VitConfig, ls_param_name: str = "gamma") -> None:
        super().__init__()
        self.cfg = cfg
        self.embed_dim = cfg.dim
        D = cfg.dim
        self.patch_embed = _PatchEmbed(D)
        if cfg.class_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        if cfg.reg_tokens:
            self.reg_token = nn.Parameter(torch.zeros(1, cfg.reg_tokens, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, NUM_PATCHES, D))
        self.blocks = nn.ModuleList([Block(cfg, ls_param_name) for _ in range(cfg.depth)])
        self.norm = nn.LayerNorm(D, eps=LN_EPS)
        if cfg.attn_pool:
            self.attn_pool = _AttentionPool(cfg)
        self.requires_grad_(False)
        self.eval()
        self._packed: Optional[_PackedTower] = None
        self.ln_folded = os.environ.get("BLB_LN_EXPLICIT") is None   # see _PackedTower; set before the first forward
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # -- packing ------------------------------------------------------------------------------------
    def invalidate_packed(self) -> None:
        self._packed = None

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() move the masters → re-pack lazily
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def packed(self) -> "_PackedTower":
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError("bridgelang_b200 towers run on CUDA only (no CPU fallback): call .cuda() first")
        if self._packed is None or self._packed.device != dev:
            self._packed = _PackedTower(self)
        return self._packed

    def workspace(self, nbytes: int) -> torch.Tensor:
        return ops.shared_workspace(self.pos_embed.device, nbytes)

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_into(self, pixels: torch.Tensor, out: torch.Tensor, col_off: int) -> None:
        """Write this tower's patch tokens into out[:, :, col_off:col_off+D] (the fused concat buffer)."""
        lib = _lib.load()
        pk = self.packed()
        B = pixels.shape[0]
        px = _as_pixels(pixels)
        if B == 0:                      # an empty shard (more ranks than images): nothing to launch
            return
        need = lib.blb_vit_workspace_bytes(C.byref(pk.struct), B)
        with ops.on_device(px, out, self.pos_embed):   # kernels, TMA maps and the stream belong to the tensors' GPU
            ws = self.workspace(need)
            _lib.check(lib.blb_vit_tower_forward(C.byref(pk.struct), px.data_ptr(), B, out.data_ptr(), out.stride(-2),
                                                 col_off, ws.data_ptr(), ws.numel(),
                                                 torch.cuda.current_stream().cuda_stream), "vit_tower_forward")

    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        B = pixels.shape[0]
        out = torch.empty((B, NUM_PATCHES, self.cfg.dim), dtype=torch.bfloat16, device=pixels.device)
        self.forward_into(pixels, out, 0)
        return out

    def get_intermediate_layers(self, x: torch.Tensor, n=None, **_: Any) -> Tuple[torch.Tensor]:
        """timm signature used at dinosiglip_vit.py:61-66; only n = {depth-2} (the reference's choice) exists."""
        want = {len(self.blocks) - 2}
        if n is not None and (set(n) if not isinstance(n, int) else {len(self.blocks) - n}) != want:
            raise ValueError("only the second-to-last block output is available on the native path")
        return (self.forward(x),)


def _as_pixels(pixels: torch.Tensor) -> torch.Tensor:
    if not pixels.is_cuda:
        raise RuntimeError("pixel_values must be CUDA tensors (no CPU fallback)")
    if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, 224, 224):
        raise ValueError(f"expected pixel_values [B,3,224,224], got {tuple(pixels.shape)}")
    return pixels.to(torch.bfloat16).contiguous()


class _PackedTower:
    """Device-resident operands in the layout the kernels want: bf16 GEMM weights (fc1/fc2 zero-padded to
    hidden_pad, patch weight flattened to [D, 592]), fp32 biases / LayerNorm / LayerScale / pos-embed / prefix
    rows, plus the ctypes descriptor structs.  fp32 masters are rounded to bf16 exactly once, here."""

    def __init__(self, vit: VisionTransformer) -> None:
        cfg = vit.cfg
        dev = vit.pos_embed.device
        self.device = dev
        self.keep: List[torch.Tensor] = []
        D, Hm, Hp = cfg.dim, cfg.mlp_hidden, cfg.hidden_pad

        def f32(t: torch.Tensor) -> torch.Tensor:
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            self.keep.append(t)
            return t

        def b16(t: torch.Tensor) -> torch.Tensor:
            t = t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
            self.keep.append(t)
            return t

        pw = torch.zeros((D, PATCH_LDK), dtype=torch.float32, device=dev)
        pw[:, :PATCH_K] = vit.patch_embed.proj.weight.detach().reshape(D, PATCH_K).float()   # (c, kh, kw) order
        prefix = []
        if cfg.class_token:
            prefix.append(vit.cls_token.detach().reshape(1, D))
        if cfg.reg_tokens:
            prefix.append(vit.reg_token.detach().reshape(cfg.reg_tokens, D))
        self.prefix = f32(torch.cat(prefix, dim=0)) if prefix else None

        # norm1 → attn.qkv and norm2 → mlp.fc1 folded (default): LN(x)·Wᵀ + b = rstd·(x·W'ᵀ − mean·colsum(W')) + b'
        # with W' = W·diag(ln_w) rounded to bf16 once, colsum over the ROUNDED W' (so the mean term cancels exactly
        # what the tensor cores accumulate) and b' = b + W·ln_b in fp32.  BLB_LN_EXPLICIT=1 keeps the LayerNorm kernel.
        self.ln_folded = vit.ln_folded

        def fold(w: torch.Tensor, b: torch.Tensor, ln: nn.LayerNorm):
            w32, lw, lb = w.float(), ln.weight.detach().float().to(dev), ln.bias.detach().float().to(dev)
            wf = b16(w32 * lw[None, :])
            return wf, f32(b.float() + w32 @ lb), f32(wf.float().sum(dim=1))

        n_blocks = cfg.n_needed_blocks
        self.blocks = (_lib.BlockWeights * n_blocks)()
        for i in range(n_blocks):
            blk: Block = vit.blocks[i]
            bw = self.blocks[i]
            fc1_w = torch.zeros((Hp, D), dtype=torch.float32, device=dev)
            fc1_w[:Hm] = blk.mlp.fc1.weight.detach().float()
            fc1_b = torch.zeros((Hp,), dtype=torch.float32, device=dev)
            fc1_b[:Hm] = blk.mlp.fc1.bias.detach().float()
            fc2_w = torch.zeros((D, Hp), dtype=torch.float32, device=dev)
            fc2_w[:, :Hm] = blk.mlp.fc2.weight.detach().float()
            bw.proj_w, bw.proj_b = b16(blk.attn.proj.weight).data_ptr(), f32(blk.attn.proj.bias).data_ptr()
            if self.ln_folded:
                qw, qb, qs = fold(blk.attn.qkv.weight.detach().to(dev), blk.attn.qkv.bias.detach().to(dev), blk.norm1)
                fw, fb, fs = fold(fc1_w, fc1_b, blk.norm2)
                bw.qkv_w, bw.qkv_b, bw.qkv_colsum = qw.data_ptr(), qb.data_ptr(), qs.data_ptr()
                bw.fc1_w, bw.fc1_b, bw.fc1_colsum = fw.data_ptr(), fb.data_ptr(), fs.data_ptr()
                bw.ln1_w = bw.ln1_b = bw.ln2_w = bw.ln2_b = None
            else:
                bw.ln1_w, bw.ln1_b = f32(blk.norm1.weight).data_ptr(), f32(blk.norm1.bias).data_ptr()
                bw.qkv_w, bw.qkv_b = b16(blk.attn.qkv.weight).data_ptr(), f32(blk.attn.qkv.bias).data_ptr()
                bw.ln2_w, bw.ln2_b = f32(blk.norm2.weight).data_ptr(), f32(blk.norm2.bias).data_ptr()
                bw.fc1_w, bw.fc1_b = b16(fc1_w).data_ptr(), f32(fc1_b).data_ptr()
                bw.qkv_colsum = bw.fc1_colsum = None
            bw.fc2_w, bw.fc2_b = b16(fc2_w).data_ptr(), f32(blk.mlp.fc2.bias).data_ptr()
            if cfg.layer_scale:
                bw.ls1, bw.ls2 = f32(blk.ls1.value).data_ptr(), f32(blk.ls2.value).data_ptr()
            else:
                bw.ls1, bw.ls2 = None, None

        s = _lib.VitWeights()
        s.dim, s.heads, s.head_dim, s.hidden_pad = D, cfg.heads, cfg.head_dim, Hp
        s.n_prefix, s.n_blocks, s.patch_ldk, s.ln_eps = cfg.n_prefix, n_blocks, PATCH_LDK, LN_EPS
        s.patch_w = b16(pw).data_ptr()
        s.patch_b = f32(vit.patch_embed.proj.bias).data_ptr()
        s.pos_embed = f32(vit.pos_embed.reshape(NUM_PATCHES, D)).data_ptr()
        s.prefix = self.prefix.data_ptr() if self.prefix is not None else None
        s.blocks_host = C.cast(self.blocks, C.POINTER(_lib.BlockWeights))
        s.ln_folded = 1 if self.ln_folded else 0
        s.hidden = Hm
        self.struct = s


# ======================================================================================================
# Image transforms (host side, upstream of the measured path: SURVEY.md §8 a1)
# ======================================================================================================
@dataclass
class LetterboxPad:
    """base_vision.py:42-51"""
    padding_fill_value: Tuple[int, int, int]

    def __call__(self, image):
        import torchvision.transforms.functional as TVF
        (w, h), max_wh = image.size, max(image.size)
        horizontal_pad, vertical_pad = int((max_wh - w) / 2), int((max_wh - h) / 2)
        padding = (horizontal_pad, vertical_pad, horizontal_pad, vertical_pad)
        return TVF.pad(image, padding, fill=self.padding_fill_value, padding_mode="constant")


@dataclass
class DinoSigLIPImageTransform:
    """dinosiglip_vit.py:33-40"""
    dino_image_transform: Callable
    siglip_image_transform: Callable
    is_prismatic: bool = True

    def __call__(self, img, **kwargs: str) -> Dict[str, torch.Tensor]:
        return {"dino": self.dino_image_transform(img, **kwargs), "siglip": self.siglip_image_transform(img, **kwargs)}


def make_preprocess_lut(device="cuda") -> torch.Tensor:
    """bf16 [2 towers (dino, siglip), 3 channels, 256]: ToTensor (x/255) → Normalize ((t−mean)/std, fp32) → bf16 cast,
    evaluated with the reference transform's own torch ops on every uint8 value (processing_prismatic.py:128-145,
    dinosiglip_vit.py:33-40) — the table `blb_preprocess_u8` gathers from, hence bit-identical to the host path."""
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)                       # ToTensor
    rows = []
    for mean, std in ((DINO_MEAN, DINO_STD), (SIGLIP_MEAN, SIGLIP_STD)):
        m = torch.as_tensor(mean, dtype=torch.float32)[:, None]
        sd = torch.as_tensor(std, dtype=torch.float32)[:, None]
        rows.append(v[None, :].repeat(3, 1).sub_(m).div_(sd))                                 # Normalize
    return torch.stack(rows).to(torch.bfloat16).to(device).contiguous()


def _make_transform(strategy: str, size: int, mean, std, *, crop_resize: int):
    """What timm.data.create_transform(is_training=False) yields for these checkpoints (bicubic Resize →
    CenterCrop → ToTensor → Normalize), modified per `image_resize_strategy` as dinosiglip_vit.py:82-134 does."""
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode

    tail = [T.CenterCrop(size), T.ToTensor(), T.Normalize(mean=torch.tensor(mean), std=torch.tensor(std))]
    if strategy == "resize-naive":
        return T.Compose([T.Resize((size, size), interpolation=InterpolationMode.BICUBIC), *tail])
    if strategy == "resize-crop":
        return T.Compose([T.Resize(crop_resize, interpolation=InterpolationMode.BICUBIC), *tail])
    if strategy == "letterbox":
        fill = tuple(int(x * 255) for x in mean)
        return T.Compose([LetterboxPad(fill), T.Resize(crop_resize, interpolation=InterpolationMode.BICUBIC), *tail])
    raise ValueError(f"Image Resize Strategy `{strategy}` is not supported!")


def letterbox_pad_transform(image, padding_fill_value: Tuple[int, int, int]):
    """processing_prismatic.py:22-28"""
    import torchvision.transforms.functional as TVF
    (w, h), max_wh = image.size, max(image.size)
    horizontal_pad, vertical_pad = int((max_wh - w) / 2), int((max_wh - h) / 2)
    padding = (horizontal_pad, vertical_pad, horizontal_pad, vertical_pad)
    return TVF.pad(image, padding, fill=padding_fill_value, padding_mode="constant")


class PrismaticImageProcessor:
    """HF twin of the image transform, prismatic/extern/hf/processing_prismatic.py:31-170: same constructor arguments
    and attributes (`tvf_resize_params`, `tvf_crop_params`, `tvf_normalize_params`, `tvf_do_letterbox`, ...), same
    `apply_transform(img) -> [3·n_towers, H, W]` (towers channel-stacked, dino first) and `preprocess(images)`.
    The reference derives the per-tower (Resize, CenterCrop, ToTensor, Normalize) parameters by parsing a transform
    that timm builds; timm is not a dependency here, and the parsed result is fully determined by the constructor
    arguments (crop_pct = 1.0 → Resize(size = input_size[-1], bicubic, antialias), CenterCrop(input_size[-2:])), so the
    parameters are written down directly.  `preprocess_to_device` is the SURVEY §8f.2 fast path: one host resize to a
    uint8 frame, H2D of 3 bytes per pixel, ToTensor + every tower's Normalize + bf16 cast in one device kernel."""

    model_input_names = ["pixel_values"]

    def __init__(self, use_fused_vision_backbone: bool = False, image_resize_strategy: str = "letterbox",
                 input_sizes: Optional[List[Tuple[int, int, int]]] = None, interpolations: Optional[List[str]] = None,
                 means: Optional[List[Tuple[float, float, float]]] = None,
                 stds: Optional[List[Tuple[float, float, float]]] = None, **kwargs: str) -> None:
        from torchvision.transforms import InterpolationMode
        import torchvision.transforms.functional as TVF
        self.use_fused_vision_backbone = use_fused_vision_backbone
        self.image_resize_strategy = image_resize_strategy
        input_sizes = [(3, 224, 224)] if input_sizes is None else input_sizes
        means = [(0.5, 0.5, 0.5)] if means is None else means
        stds = [(0.5, 0.5, 0.5)] if stds is None else stds
        interpolations = ["bicubic"] * len(input_sizes) if interpolations is None else interpolations
        self.input_sizes, self.interpolations, self.means, self.stds = input_sizes, interpolations, means, stds
        self.tvf_resize_params, self.tvf_crop_params, self.tvf_normalize_params = [], [], []
        self.tvf_do_letterbox, self.tvf_letterbox_fill = False, None
        modes = {"bicubic": InterpolationMode.BICUBIC, "bilinear": InterpolationMode.BILINEAR,
                 "nearest": InterpolationMode.NEAREST}
        for idx in range(len(input_sizes)):
            size = self.input_sizes[idx][-1]
            self.tvf_resize_params.append({"size": size, "interpolation": TVF.pil_modes_mapping[modes[interpolations[idx]]],
                                           "max_size": None, "antialias": True})
            self.tvf_crop_params.append({"output_size": tuple(self.input_sizes[idx][-2:])})
            self.tvf_normalize_params.append({"mean": torch.tensor(self.means[idx]).float().numpy().tolist(),
                                              "std": torch.tensor(self.stds[idx]).float().numpy().tolist(),
                                              "inplace": False})
            self.tvf_do_letterbox, self.tvf_letterbox_fill = False, None
            if self.image_resize_strategy == "resize-naive":
                self.tvf_resize_params[idx]["size"] = (size, size)
            elif self.image_resize_strategy == "letterbox":
                self.tvf_do_letterbox, self.tvf_letterbox_fill = True, tuple([int(x * 255) for x in self.means[idx]])
            elif self.image_resize_strategy == "resize-crop":
                pass
            else:
                raise ValueError(f"Image resize strategy `{self.image_resize_strategy}` is not supported!")
        self._lut = None

    def _resized(self, img, idx: int):
        import torchvision.transforms.functional as TVF
        return TVF.center_crop(TVF.resize(img, **self.tvf_resize_params[idx]), **self.tvf_crop_params[idx])

    def apply_transform(self, img) -> torch.Tensor:
        """processing_prismatic.py:128-145"""
        import torchvision.transforms.functional as TVF
        if self.tvf_do_letterbox:
            img = letterbox_pad_transform(img, self.tvf_letterbox_fill)
        imgs_t = []
        for idx in range(len(self.input_sizes)):
            img_idx_t = TVF.to_tensor(self._resized(img, idx))
            imgs_t.append(TVF.normalize(img_idx_t, **self.tvf_normalize_params[idx]))
        return torch.vstack(imgs_t)

    def preprocess(self, images, return_tensors: Optional[str] = None, **_: str) -> Dict[str, Any]:
        """Returns {"pixel_values": ...} (a plain dict standing in for transformers' BatchFeature): float32 torch tensor
        for return_tensors="pt", NumPy otherwise — as processing_prismatic.py:147-167."""
        if not isinstance(images, list):
            images = [images]
        pixel_values = torch.stack([self.apply_transform(img.convert("RGB")) for img in images])
        return {"pixel_values": pixel_values.float() if return_tensors == "pt" else pixel_values.float().numpy()}

    def __call__(self, images, **kwargs):
        return self.preprocess(images, **kwargs)

    def preprocess_to_device(self, images, device="cuda") -> torch.Tensor:
        """[B, 6, 224, 224] bf16 on the device, bit-identical to `preprocess(...)["pixel_values"].to(device, bf16)` for the
        fused 224 px DINOv2 + SigLIP configuration (both towers resize identically, so one uint8 frame feeds both)."""
        import numpy as np
        if not (self.use_fused_vision_backbone and len(self.input_sizes) == 2
                and all(tuple(sz) == (3, 224, 224) for sz in self.input_sizes)
                and self.tvf_resize_params[0] == self.tvf_resize_params[1]
                and [tuple(m) for m in self.means] == [tuple(DINO_MEAN), tuple(SIGLIP_MEAN)]
                and [tuple(sd) for sd in self.stds] == [tuple(DINO_STD), tuple(SIGLIP_STD)]):
            raise ValueError("preprocess_to_device covers the fused 224 px DINOv2 + SigLIP processor configuration")
        if not isinstance(images, list):
            images = [images]
        frames = []
        for img in images:
            img = img.convert("RGB")
            if self.tvf_do_letterbox:
                img = letterbox_pad_transform(img, self.tvf_letterbox_fill)
            frames.append(np.asarray(self._resized(img, 0), dtype=np.uint8))
        u8 = torch.from_numpy(np.stack(frames)).pin_memory().to(device, non_blocking=True)
        if self._lut is None or self._lut.device != u8.device:
            self._lut = make_preprocess_lut(u8.device)
        dino, siglip = ops.preprocess_u8(u8, self._lut)
        return torch.cat([dino, siglip], dim=1)


# ======================================================================================================
# Reference-facing backbones
# ======================================================================================================
class VisionBackbone(nn.Module, ABC):
    """base_vision.py:54-90, unchanged contract."""

    def __init__(self, vision_backbone_id: str, image_resize_strategy: str, default_image_size: int = 224) -> None:
        super().__init__()
        self.identifier: str = vision_backbone_id
        self.image_resize_strategy: str = image_resize_strategy
        self.default_image_size: int = default_image_size
        self.featurizer: nn.Module = None
        self.image_transform = None

    def get_image_transform(self):
        return self.image_transform

    @abstractmethod
    def get_fsdp_wrapping_policy(self) -> Callable: ...

    @abstractmethod
    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor: ...

    @property
    @abstractmethod
    def default_image_resolution(self) -> Tuple[int, int, int]: ...

    @property
    @abstractmethod
    def embed_dim(self) -> int: ...

    @property
    @abstractmethod
    def num_patches(self) -> int: ...

    @property
    @abstractmethod
    def half_precision_dtype(self) -> torch.dtype: ...


def _vit_fsdp_policy() -> Callable:
    from torch.distributed.fsdp.wrap import _module_wrap_policy, _or_policy, transformer_auto_wrap_policy
    vit_wrap_policy = partial(_module_wrap_policy, module_classes={VisionTransformer})
    transformer_block_policy = partial(transformer_auto_wrap_policy, transformer_layer_cls={Block})
    return partial(_or_policy, policies=[vit_wrap_policy, transformer_block_policy])


class DinoSigLIPViTBackbone(VisionBackbone):
    """dinosiglip_vit.py:43-163.  `forward({"dino": [B,3,224,224], "siglip": [B,3,224,224]}) -> [B,256,2176]`
    (cols 0-1023 DINOv2, 1024-2175 SigLIP).  Each tower's last needed fc2 epilogue writes its column slice of the
    output directly, so the reference's torch.cat copy does not exist."""

    def __init__(self, vision_backbone_id: str, image_resize_strategy: str, default_image_size: int = 224) -> None:
        super().__init__(vision_backbone_id, image_resize_strategy, default_image_size=default_image_size)
        if vision_backbone_id not in DINOSigLIP_VISION_BACKBONES:
            raise ValueError(f"Vision Backbone `{vision_backbone_id}` is not supported on the B200-native path!")
        if default_image_size != 224:
            raise ValueError("the B200-native towers are built for 224 px inputs")
        cfgs = DINOSigLIP_VISION_BACKBONES[vision_backbone_id]
        self.dino_timm_path_or_url = cfgs["dino"].timm_id
        self.siglip_timm_path_or_url = cfgs["siglip"].timm_id
        self.dino_featurizer = VisionTransformer(cfgs["dino"])
        self.siglip_featurizer = VisionTransformer(cfgs["siglip"])
        self.dtype = torch.bfloat16
        self.dino_data_cfg = {"input_size": (3, 224, 224), "interpolation": "bicubic", "mean": DINO_MEAN,
                              "std": DINO_STD, "crop_pct": 1.0, "crop_mode": "center"}
        self.siglip_data_cfg = {"input_size": (3, 224, 224), "interpolation": "bicubic", "mean": SIGLIP_MEAN,
                                "std": SIGLIP_STD, "crop_pct": 0.9, "crop_mode": "center"}
        self.image_transform = DinoSigLIPImageTransform(
            _make_transform(image_resize_strategy, 224, DINO_MEAN, DINO_STD, crop_resize=224),
            _make_transform(image_resize_strategy, 224, SIGLIP_MEAN, SIGLIP_STD, crop_resize=224),
        )

    def get_fsdp_wrapping_policy(self) -> Callable:
        return _vit_fsdp_policy()

    def forward(self, pixel_values: Dict[str, torch.Tensor]) -> torch.Tensor:
        dino_px, siglip_px = pixel_values["dino"], pixel_values["siglip"]
        B = dino_px.shape[0]
        out = torch.empty((B, NUM_PATCHES, self.embed_dim), dtype=torch.bfloat16, device=dino_px.device)
        self.dino_featurizer.forward_into(dino_px, out, 0)
        self.siglip_featurizer.forward_into(siglip_px, out, self.dino_featurizer.embed_dim)
        return out

    def preprocess_uint8(self, frames: torch.Tensor) -> Dict[str, torch.Tensor]:
        """already-resized uint8 frames [B,224,224,3] (HWC, on the GPU) → the `pixel_values` dict in bf16,
        one device kernel for both towers instead of two host-side ToTensor+Normalize passes and a 4x larger H2D copy."""
        if getattr(self, "_lut", None) is None or self._lut.device != frames.device:
            self._lut = make_preprocess_lut(frames.device)
        dino, siglip = ops.preprocess_u8(frames, self._lut)
        return {"dino": dino, "siglip": siglip}

    @property
    def default_image_resolution(self) -> Tuple[int, int, int]:
        return self.dino_data_cfg["input_size"]

    @property
    def embed_dim(self) -> int:
        return self.dino_featurizer.embed_dim + self.siglip_featurizer.embed_dim

    @property
    def num_patches(self) -> int:
        return NUM_PATCHES

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return torch.bfloat16


class _SingleTowerBackbone(VisionBackbone):
    """TimmViTBackbone (base_vision.py:94-207) restricted to the two towers this path owns (BASELINE config 4)."""

    REGISTRY: Dict[str, VitConfig] = {}
    MEAN: Tuple[float, float, float] = (0.5, 0.5, 0.5)
    STD: Tuple[float, float, float] = (0.5, 0.5, 0.5)

    def __init__(self, vision_backbone_id: str, image_resize_strategy: str, default_image_size: int = 224) -> None:
        super().__init__(vision_backbone_id, image_resize_strategy, default_image_size=default_image_size)
        if vision_backbone_id not in self.REGISTRY:
            raise ValueError(f"Vision Backbone `{vision_backbone_id}` is not supported on the B200-native path!")
        cfg = self.REGISTRY[vision_backbone_id]
        self.timm_path_or_url = cfg.timm_id
        self.dtype = torch.bfloat16
        self.featurizer = VisionTransformer(cfg)
        self.data_cfg = {"input_size": (3, 224, 224), "mean": self.MEAN, "std": self.STD}
        self.image_transform = _make_transform(image_resize_strategy, 224, self.MEAN, self.STD, crop_resize=224)

    def get_fsdp_wrapping_policy(self) -> Callable:
        return _vit_fsdp_policy()

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        return self.featurizer(pixel_values)

    @property
    def default_image_resolution(self) -> Tuple[int, int, int]:
        return self.data_cfg["input_size"]

    @property
    def embed_dim(self) -> int:
        return self.featurizer.embed_dim

    @property
    def num_patches(self) -> int:
        return NUM_PATCHES

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return self.dtype


class SigLIPViTBackbone(_SingleTowerBackbone):
    """siglip_vit.py:8-24 (`siglip-vit-so400m`)."""
    REGISTRY = SIGLIP_VISION_BACKBONES
    MEAN, STD = SIGLIP_MEAN, SIGLIP_STD


class DinoV2ViTBackbone(_SingleTowerBackbone):
    """dinov2_vit.py:9-19 (`dinov2-vit-l`)."""
    REGISTRY = DINOv2_VISION_BACKBONES
    MEAN, STD = DINO_MEAN, DINO_STD


class PrismaticVisionBackbone(nn.Module):
    """HF twin, modeling_prismatic.py:63-123: attrs `featurizer`, `fused_featurizer`, `embed_dim`;
    `forward(pixel_values[B,6,224,224])` (dino channels first); LayerScale parameter named `scale_factor`."""

    def __init__(self, use_fused_vision_backbone: bool, image_sizes: List[int], timm_model_ids: List[str],
                 timm_override_act_layers: List[Optional[str]]) -> None:
        super().__init__()
        self.use_fused_vision_backbone = use_fused_vision_backbone
        assert len(timm_model_ids) <= 2, "Prismatic models only support up to 2 (fused) vision backbones!"
        if any(s != 224 for s in image_sizes) or any(a is not None for a in timm_override_act_layers):
            raise ValueError("the B200-native towers support 224 px inputs and timm's default (erf GELU) activation")
        for tid in timm_model_ids:
            if tid not in TIMM_ID_TO_CONFIG:
                raise ValueError(f"timm model `{tid}` is not supported on the B200-native path")
        self.featurizer = VisionTransformer(TIMM_ID_TO_CONFIG[timm_model_ids[0]], ls_param_name="scale_factor")
        self.embed_dim = self.featurizer.embed_dim
        if self.use_fused_vision_backbone:
            self.fused_featurizer = VisionTransformer(TIMM_ID_TO_CONFIG[timm_model_ids[1]],
                                                      ls_param_name="scale_factor")
            self.embed_dim += self.fused_featurizer.embed_dim

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        if not self.use_fused_vision_backbone:
            return self.featurizer(pixel_values)
        img, img_fused = torch.split(pixel_values, [3, 3], dim=1)
        B = pixel_values.shape[0]
        out = torch.empty((B, NUM_PATCHES, self.embed_dim), dtype=torch.bfloat16, device=pixel_values.device)
        self.featurizer.forward_into(img, out, 0)
        self.fused_featurizer.forward_into(img_fused, out, self.featurizer.embed_dim)
        return out
