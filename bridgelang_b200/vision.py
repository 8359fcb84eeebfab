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
from .config import (CLIP_L14_336, DINOV2_L14_REG4, DINOV2_L14_REG4_336, DINOV2_L14_REG4_384, LN_EPS, NUM_PATCHES, PATCH,
                     PATCH_K, PATCH_LDK, SIGLIP_SO400M_14, SIGLIP_SO400M_14_384, VitConfig)
from .weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD

# Registry =>> same ids as dinosiglip_vit.py:21-30 / dinoclip_vit.py:22-27 / siglip_vit.py:8-13 / dinov2_vit.py:9-10 /
# clip_vit.py:8-12, keyed additionally by the `default_image_size` materialize.py:35-49 passes
DINOSigLIP_VISION_BACKBONES = {
    "dinosiglip-vit-so-224px": {"dino": DINOV2_L14_REG4, "siglip": SIGLIP_SO400M_14},
    "dinosiglip-vit-so-384px": {"dino": DINOV2_L14_REG4_384, "siglip": SIGLIP_SO400M_14_384},
}
DINOCLIP_VISION_BACKBONES = {
    "dinoclip-vit-l-336px": {"dino": DINOV2_L14_REG4_336, "clip": CLIP_L14_336},
}
SIGLIP_VISION_BACKBONES = {"siglip-vit-so400m": SIGLIP_SO400M_14, "siglip-vit-so400m-384px": SIGLIP_SO400M_14_384}
DINOv2_VISION_BACKBONES = {"dinov2-vit-l": DINOV2_L14_REG4}
CLIP_VISION_BACKBONES = {"clip-vit-l-336px": CLIP_L14_336}
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
    def __init__(self, dim: int, bias: bool = True) -> None:
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH, bias=bias)


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
        self.patch_embed = _PatchEmbed(D, bias=cfg.patch_bias)     # timm: bias = not pre_norm
        if cfg.class_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        if cfg.reg_tokens:
            self.reg_token = nn.Parameter(torch.zeros(1, cfg.reg_tokens, D))
        # timm: num_patches rows when no_embed_class, else num_patches + num_prefix_tokens (the CLIP checkpoints)
        self.pos_embed = nn.Parameter(torch.zeros(1, cfg.num_patches + (0 if cfg.no_embed_class else cfg.n_prefix), D))
        if cfg.pre_norm:
            self.norm_pre = nn.LayerNorm(D, eps=LN_EPS)
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
        px = _as_pixels(pixels, self.cfg.img_size)
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
        out = torch.empty((B, self.cfg.num_patches, self.cfg.dim), dtype=torch.bfloat16, device=pixels.device)
        self.forward_into(pixels, out, 0)
        return out

    @torch.no_grad()
    def forward_uint8_into(self, frames: torch.Tensor, out: torch.Tensor, col_off: int) -> None:
        """uint8 HWC frames [B, img, img, 3] (already resized; on the GPU) → this tower's column slice of `out`.
        ToTensor + this tower's Normalize are folded into the patch-embed weights: the frame goes
        through one permutation kernel into the TMA-loaded A operand of the patch-embed GEMM."""
        lib = _lib.load()
        pk = self.packed()
        fr = _as_frames(frames, self.cfg.img_size)
        B = fr.shape[0]
        if B == 0:
            return
        need = lib.blb_vit_workspace_bytes(C.byref(pk.struct), B)
        with ops.on_device(fr, out, self.pos_embed):
            ws = self.workspace(need)
            _lib.check(lib.blb_vit_tower_forward_u8(C.byref(pk.struct), fr.data_ptr(), B, out.data_ptr(), out.stride(-2),
                                                    col_off, ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream().cuda_stream), "vit_tower_forward_u8")

    def forward_uint8(self, frames: torch.Tensor) -> torch.Tensor:
        out = torch.empty((frames.shape[0], self.cfg.num_patches, self.cfg.dim), dtype=torch.bfloat16,
                          device=frames.device)
        self.forward_uint8_into(frames, out, 0)
        return out

    def get_intermediate_layers(self, x: torch.Tensor, n=None, **_: Any) -> Tuple[torch.Tensor]:
        """timm signature used at dinosiglip_vit.py:61-66; only n = {depth-2} (the reference's choice) exists."""
        want = {len(self.blocks) - 2}
        if n is not None and (set(n) if not isinstance(n, int) else {len(self.blocks) - n}) != want:
            raise ValueError("only the second-to-last block output is available on the native path")
        return (self.forward(x),)


def _as_pixels(pixels: torch.Tensor, img: int = 224) -> torch.Tensor:
    if not pixels.is_cuda:
        raise RuntimeError("pixel_values must be CUDA tensors (no CPU fallback)")
    if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, img, img):
        raise ValueError(f"expected pixel_values [B,3,{img},{img}], got {tuple(pixels.shape)}")
    return pixels.to(torch.bfloat16).contiguous()


def _as_frames(frames: torch.Tensor, img: int = 224) -> torch.Tensor:
    if not frames.is_cuda:
        raise RuntimeError("uint8 frames must be CUDA tensors (no CPU fallback)")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or tuple(frames.shape[1:]) != (img, img, 3):
        raise ValueError(f"expected uint8 frames [B,{img},{img},3] (HWC), got {frames.dtype} {tuple(frames.shape)}")
    return frames.contiguous()


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

        conv_w = vit.patch_embed.proj.weight.detach().to(dev).float()                          # [D, 3, 14, 14]
        conv_b = (vit.patch_embed.proj.bias.detach().to(dev).float() if cfg.patch_bias
                  else torch.zeros(D, dtype=torch.float32, device=dev))
        pw = torch.zeros((D, PATCH_LDK), dtype=torch.float32, device=dev)
        pw[:, :PATCH_K] = conv_w.reshape(D, PATCH_K)                                           # (c, kh, kw) order
        # uint8 entry: ToTensor + Normalize folded in — conv((u/255 − mean_c)/std_c) =
        # Σ W/(255·std_c)·u + (b − Σ W·mean_c/std_c) — and K reordered to (kh, kw, c), the order of an HWC patch row
        mean = torch.tensor(cfg.mean, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
        std = torch.tensor(cfg.std, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
        pw8 = torch.zeros((D, PATCH_LDK), dtype=torch.float32, device=dev)
        pw8[:, :PATCH_K] = (conv_w / (255.0 * std)).permute(0, 2, 3, 1).reshape(D, PATCH_K)
        pb8 = conv_b - (conv_w * (mean / std)).sum(dim=(1, 2, 3))
        # prefix rows (cls, reg): input independent.  With no_embed_class=False (CLIP) pos_embed also has rows for them:
        # fold those into the prefix rows and hand the kernels the patch rows only
        pos = vit.pos_embed.detach().to(dev).float().reshape(-1, D)
        pos_prefix = None
        if not cfg.no_embed_class and cfg.n_prefix:
            pos_prefix, pos = pos[:cfg.n_prefix], pos[cfg.n_prefix:]
        prefix = []
        if cfg.class_token:
            prefix.append(vit.cls_token.detach().to(dev).float().reshape(1, D))
        if cfg.reg_tokens:
            prefix.append(vit.reg_token.detach().to(dev).float().reshape(cfg.reg_tokens, D))
        if prefix:
            rows = torch.cat(prefix, dim=0)
            self.prefix = f32(rows + pos_prefix if pos_prefix is not None else rows)
        else:
            self.prefix = None

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
        s.patch_b = f32(conv_b).data_ptr() if cfg.patch_bias else None
        s.pos_embed = f32(pos.reshape(cfg.num_patches, D)).data_ptr()
        s.prefix = self.prefix.data_ptr() if self.prefix is not None else None
        s.grid, s.img_size = cfg.grid, cfg.img_size
        s.act = {"gelu": 0, "quick_gelu": 1}[cfg.act]
        if cfg.pre_norm:
            s.norm_pre_w, s.norm_pre_b = f32(vit.norm_pre.weight).data_ptr(), f32(vit.norm_pre.bias).data_ptr()
        s.patch_w_u8, s.patch_b_u8 = b16(pw8).data_ptr(), f32(pb8).data_ptr()
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

    def _check_fused_224(self) -> None:
        if not (self.use_fused_vision_backbone and len(self.input_sizes) == 2
                and all(tuple(sz) == (3, 224, 224) for sz in self.input_sizes)
                and self.tvf_resize_params[0] == self.tvf_resize_params[1]
                and [tuple(m) for m in self.means] == [tuple(DINO_MEAN), tuple(SIGLIP_MEAN)]
                and [tuple(sd) for sd in self.stds] == [tuple(DINO_STD), tuple(SIGLIP_STD)]):
            raise ValueError("the device preprocessing path covers the fused 224 px DINOv2 + SigLIP processor configuration")

    def frames_to_device(self, images, device="cuda", device_resize: bool = True) -> torch.Tensor:
        """PIL images → uint8 HWC frames [B,224,224,3] on the device, byte-identical to what the host transform feeds
        ToTensor (letterbox → bicubic Resize → CenterCrop).  device_resize=True: only the RAW frame crosses PCIe and the
        antialiased bicubic resize runs on the GPU (`ops.resize_u8`, bit-exact with PIL); images of different sizes
        are resized one by one.  The result feeds `VisualPrefixEncoder.forward_uint8`."""
        import numpy as np
        self._check_fused_224()
        if not isinstance(images, list):
            images = [images]
        out = []
        for img in images:
            img = img.convert("RGB")
            if self.tvf_do_letterbox:
                img = letterbox_pad_transform(img, self.tvf_letterbox_fill)
            if not device_resize:
                out.append(torch.from_numpy(np.array(self._resized(img, 0), dtype=np.uint8)).to(device))
                continue
            raw = torch.from_numpy(np.array(img, dtype=np.uint8)).to(device, non_blocking=True)[None].contiguous()
            size = self.tvf_resize_params[0]["size"]
            H, W = raw.shape[1], raw.shape[2]
            if isinstance(size, (tuple, list)):                    # resize-naive: exact (h, w)
                Hd, Wd = int(size[0]), int(size[1])
            else:                                                  # torchvision: shorter side → size, aspect kept
                short, long = (W, H) if W <= H else (H, W)
                new_short, new_long = int(size), int(size * long / short)
                Wd, Hd = (new_short, new_long) if W <= H else (new_long, new_short)
            fr = ops.resize_u8(raw, (Hd, Wd))[0]
            ch, cw = self.tvf_crop_params[0]["output_size"]
            top, left = int(round((Hd - ch) / 2.0)), int(round((Wd - cw) / 2.0))   # torchvision center_crop
            out.append(fr[top:top + ch, left:left + cw].contiguous())
        return torch.stack(out)

    def preprocess_to_device(self, images, device="cuda", device_resize: bool = False) -> torch.Tensor:
        """[B, 6, 224, 224] bf16 on the device, bit-identical to `preprocess(...)["pixel_values"].to(device, bf16)` for the
        fused 224 px DINOv2 + SigLIP configuration (both towers resize identically, so one uint8 frame feeds both):
        ToTensor + both Normalizes + the bf16 cast are one LUT kernel; with device_resize=True the bicubic resize runs
        on the GPU as well (PIL-exact), so the host only decodes the image."""
        u8 = self.frames_to_device(images, device=device, device_resize=device_resize)
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


class _FusedBackbone(VisionBackbone):
    """Two towers on the same frame, patch tokens concatenated channel-wise (dinosiglip_vit.py:142-147,
    dinoclip_vit.py:141-147).  Each tower's last needed fc2 epilogue writes its column slice of the output directly, so
    the reference's torch.cat copy does not exist."""

    KEYS: Tuple[str, str] = ("dino", "siglip")

    def _build(self, cfgs: Dict[str, VitConfig], image_resize_strategy: str, default_image_size: int) -> None:
        k0, k1 = self.KEYS
        for c in cfgs.values():
            if c.img_size != default_image_size:
                raise ValueError(f"`{self.identifier}` is registered for {c.img_size} px inputs, got "
                                 f"default_image_size={default_image_size}")
        setattr(self, f"{k0}_timm_path_or_url", cfgs[k0].timm_id)
        setattr(self, f"{k1}_timm_path_or_url", cfgs[k1].timm_id)
        setattr(self, f"{k0}_featurizer", VisionTransformer(cfgs[k0]))
        setattr(self, f"{k1}_featurizer", VisionTransformer(cfgs[k1]))
        self.dtype = torch.bfloat16
        S = default_image_size
        for k in self.KEYS:
            setattr(self, f"{k}_data_cfg", {"input_size": (3, S, S), "interpolation": "bicubic", "mean": cfgs[k].mean,
                                            "std": cfgs[k].std, "crop_pct": 1.0, "crop_mode": "center"})
        self.image_transform = self._make_pair_transform(
            _make_transform(image_resize_strategy, S, cfgs[k0].mean, cfgs[k0].std, crop_resize=S),
            _make_transform(image_resize_strategy, S, cfgs[k1].mean, cfgs[k1].std, crop_resize=S))

    def _towers(self) -> Tuple["VisionTransformer", "VisionTransformer"]:
        return getattr(self, f"{self.KEYS[0]}_featurizer"), getattr(self, f"{self.KEYS[1]}_featurizer")

    def get_fsdp_wrapping_policy(self) -> Callable:
        return _vit_fsdp_policy()

    def forward(self, pixel_values: Dict[str, torch.Tensor]) -> torch.Tensor:
        t0, t1 = self._towers()
        px0, px1 = pixel_values[self.KEYS[0]], pixel_values[self.KEYS[1]]
        B = px0.shape[0]
        out = torch.empty((B, self.num_patches, self.embed_dim), dtype=torch.bfloat16, device=px0.device)
        t0.forward_into(px0, out, 0)
        t1.forward_into(px1, out, t0.embed_dim)
        return out

    @property
    def default_image_resolution(self) -> Tuple[int, int, int]:
        return getattr(self, f"{self.KEYS[0]}_data_cfg")["input_size"]

    @property
    def embed_dim(self) -> int:
        t0, t1 = self._towers()
        return t0.embed_dim + t1.embed_dim

    @property
    def num_patches(self) -> int:
        return self._towers()[0].cfg.num_patches

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return torch.bfloat16


class DinoSigLIPViTBackbone(_FusedBackbone):
    """dinosiglip_vit.py:43-163.  `forward({"dino": [B,3,S,S], "siglip": [B,3,S,S]}) -> [B,P,2176]` (cols 0-1023 DINOv2,
    1024-2175 SigLIP); `dinosiglip-vit-so-224px` (S 224, P 256) and `dinosiglip-vit-so-384px` (S 384, P 729)."""

    KEYS = ("dino", "siglip")

    def __init__(self, vision_backbone_id: str, image_resize_strategy: str, default_image_size: int = 224) -> None:
        super().__init__(vision_backbone_id, image_resize_strategy, default_image_size=default_image_size)
        if vision_backbone_id not in DINOSigLIP_VISION_BACKBONES:
            raise ValueError(f"Vision Backbone `{vision_backbone_id}` is not supported on the B200-native path!")
        self._build(DINOSigLIP_VISION_BACKBONES[vision_backbone_id], image_resize_strategy, default_image_size)

    @staticmethod
    def _make_pair_transform(a, b):
        return DinoSigLIPImageTransform(a, b)

    def preprocess_uint8(self, frames: torch.Tensor) -> Dict[str, torch.Tensor]:
        """already-resized uint8 frames [B,224,224,3] (HWC, on the GPU) → the `pixel_values` dict in bf16,
        one device kernel for both towers instead of two host-side ToTensor+Normalize passes and a 4x larger H2D copy."""
        if self.default_image_size != 224:
            raise ValueError("the LUT preprocessing kernel is built for 224 px frames (use the folded uint8 entry)")
        if getattr(self, "_lut", None) is None or self._lut.device != frames.device:
            self._lut = make_preprocess_lut(frames.device)
        dino, siglip = ops.preprocess_u8(frames, self._lut)
        return {"dino": dino, "siglip": siglip}


@dataclass
class DinoCLIPImageTransform:
    """dinoclip_vit.py:30-37"""
    dino_image_transform: Callable
    clip_image_transform: Callable
    is_prismatic: bool = True

    def __call__(self, img, **kwargs: str) -> Dict[str, torch.Tensor]:
        return {"dino": self.dino_image_transform(img, **kwargs), "clip": self.clip_image_transform(img, **kwargs)}


class DinoCLIPViTBackbone(_FusedBackbone):
    """dinoclip_vit.py:40-147 (`dinoclip-vit-l-336px`): DINOv2 ViT-L/14-reg4 + OpenAI CLIP ViT-L/14-336 (quick-GELU
    through `override_act_layer`, clip_vit.py:15-27), both on the 336 px frame: `forward({"dino", "clip"}) -> [B,576,2048]`."""

    KEYS = ("dino", "clip")

    def __init__(self, vision_backbone_id: str, image_resize_strategy: str, default_image_size: int = 224) -> None:
        super().__init__(vision_backbone_id, image_resize_strategy, default_image_size=default_image_size)
        if vision_backbone_id not in DINOCLIP_VISION_BACKBONES:
            raise ValueError(f"Vision Backbone `{vision_backbone_id}` is not supported on the B200-native path!")
        self._build(DINOCLIP_VISION_BACKBONES[vision_backbone_id], image_resize_strategy, default_image_size)

    @staticmethod
    def _make_pair_transform(a, b):
        return DinoCLIPImageTransform(a, b)


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
        if cfg.img_size != default_image_size:
            raise ValueError(f"`{vision_backbone_id}` is registered for {cfg.img_size} px inputs, got "
                             f"default_image_size={default_image_size}")
        self.timm_path_or_url = cfg.timm_id
        self.dtype = torch.bfloat16
        self.featurizer = VisionTransformer(cfg)
        S = cfg.img_size
        self.data_cfg = {"input_size": (3, S, S), "mean": cfg.mean, "std": cfg.std}
        self.image_transform = _make_transform(image_resize_strategy, S, cfg.mean, cfg.std, crop_resize=S)

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
        return self.featurizer.cfg.num_patches

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return self.dtype


class SigLIPViTBackbone(_SingleTowerBackbone):
    """siglip_vit.py:8-24 (`siglip-vit-so400m`, `siglip-vit-so400m-384px`)."""
    REGISTRY = SIGLIP_VISION_BACKBONES
    MEAN, STD = SIGLIP_MEAN, SIGLIP_STD


class DinoV2ViTBackbone(_SingleTowerBackbone):
    """dinov2_vit.py:9-19 (`dinov2-vit-l`)."""
    REGISTRY = DINOv2_VISION_BACKBONES
    MEAN, STD = DINO_MEAN, DINO_STD


class CLIPViTBackbone(_SingleTowerBackbone):
    """clip_vit.py:15-27 (`clip-vit-l-336px`: OpenAI weights → quick-GELU override)."""
    REGISTRY = CLIP_VISION_BACKBONES


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
        out = torch.empty((B, self.featurizer.cfg.num_patches, self.embed_dim), dtype=torch.bfloat16,
                          device=pixel_values.device)
        self.featurizer.forward_into(img, out, 0)
        self.fused_featurizer.forward_into(img_fused, out, self.featurizer.embed_dim)
        return out
