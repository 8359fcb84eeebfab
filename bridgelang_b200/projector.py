"""projector.py — B200-native mirrors of the reference's vision→LLM projectors.

  FusedMLPProjector     prismatic/util/nn_utils.py:37-53        (params `projector.{0,2,4}.{weight,bias}`)
  PrismaticProjector    prismatic/extern/hf/modeling_prismatic.py:127-158  (params `fc1/fc2/fc3`)

`forward([..., 2176]) -> [..., 4096]`: three tcgen05 GEMMs with bias(+exact-erf GELU) fused in the epilogue
(Linear → GELU → Linear → GELU → Linear).  bf16 in / bf16 out, fp32 accumulation and bias.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops


class _ProjectorBase(nn.Module):
    def _linears(self) -> Tuple[nn.Linear, nn.Linear, nn.Linear]:
        raise NotImplementedError

    def _init_native(self) -> None:
        self._packed: Optional[Tuple[_lib.ProjectorWeights, List[torch.Tensor], torch.device]] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    def invalidate_packed(self) -> None:
        self._packed = None

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def packed(self) -> _lib.ProjectorWeights:
        fc1, fc2, fc3 = self._linears()
        dev = fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("bridgelang_b200 projector runs on CUDA only (no CPU fallback): call .cuda() first")
        if self._packed is None or self._packed[2] != dev:
            keep: List[torch.Tensor] = []

            def b16(t):
                keep.append(t.detach().to(torch.bfloat16).contiguous())
                return keep[-1].data_ptr()

            def f32(t):
                keep.append(t.detach().to(torch.float32).contiguous())
                return keep[-1].data_ptr()

            s = _lib.ProjectorWeights()
            s.in_dim, s.hidden_dim, s.out_dim = fc1.in_features, fc1.out_features, fc3.out_features
            s.fc1_w, s.fc1_b = b16(fc1.weight), f32(fc1.bias)
            s.fc2_w, s.fc2_b = b16(fc2.weight), f32(fc2.bias)
            s.fc3_w, s.fc3_b = b16(fc3.weight), f32(fc3.bias)
            self._packed = (s, keep, dev)
        return self._packed[0]

    @torch.no_grad()
    def project(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, tok_in: int = 0, tok_out: int = 0,
                tok_shift: int = 0) -> torch.Tensor:
        """x [..., in_dim] → [..., out_dim].  With `out` = a preallocated inputs_embeds buffer [B, tok_out, out_dim]
        and tok_in/tok_shift set, fc3's epilogue stores each image's 256 rows at token offset `tok_shift`
        (the multimodal splice of prismatic.py:389-396 without the cat copy)."""
        if not x.is_cuda:
            raise RuntimeError("projector input must be a CUDA tensor (no CPU fallback)")
        lib = _lib.load()
        s = self.packed()
        lead = x.shape[:-1]
        x2 = x.to(torch.bfloat16).reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        rows = x2.shape[0]
        if out is None:
            out = torch.empty((*lead, s.out_dim), dtype=torch.bfloat16, device=x.device)
            ld_out = s.out_dim
        else:
            assert out.dtype == torch.bfloat16 and out.is_cuda and out.stride(-1) == 1
            ld_out = out.stride(-2)
        if rows == 0:
            return out
        need = lib.blb_projector_workspace_bytes(C.byref(s), rows)
        with ops.on_device(x2, out, self._linears()[0].weight):
            ws = ops.shared_workspace(x.device, need)
            _lib.check(lib.blb_projector_forward(C.byref(s), x2.data_ptr(), x2.stride(0), rows, out.data_ptr(), ld_out,
                                                 tok_in, tok_out, tok_shift, ws.data_ptr(), ws.numel(),
                                                 torch.cuda.current_stream().cuda_stream),
                       "projector_forward")
        return out


class FusedMLPProjector(_ProjectorBase):
    def __init__(self, fused_vision_dim: int, llm_dim: int, mlp_type: str = "fused-gelu-mlp") -> None:
        super().__init__()
        self.initial_projection_dim = fused_vision_dim * 4
        if mlp_type == "fused-gelu-mlp":
            self.projector = nn.Sequential(
                nn.Linear(fused_vision_dim, self.initial_projection_dim, bias=True),
                nn.GELU(),
                nn.Linear(self.initial_projection_dim, llm_dim, bias=True),
                nn.GELU(),
                nn.Linear(llm_dim, llm_dim, bias=True),
            )
        else:
            raise ValueError(f"Fused Projector with `{mlp_type = }` is not supported!")
        self.requires_grad_(False)
        self._init_native()

    def _linears(self):
        return self.projector[0], self.projector[2], self.projector[4]

    def forward(self, fused_img_patches: torch.Tensor) -> torch.Tensor:
        return self.project(fused_img_patches)


class PrismaticProjector(_ProjectorBase):
    def __init__(self, use_fused_vision_backbone: bool, vision_dim: int, llm_dim: int) -> None:
        super().__init__()
        if not use_fused_vision_backbone:
            raise ValueError("the B200-native projector implements the fused (3-layer) variant of this path")
        self.use_fused_vision_backbone = use_fused_vision_backbone
        self.vision_dim, self.llm_dim = vision_dim, llm_dim
        initial_projection_dim = 4 * vision_dim
        self.fc1 = nn.Linear(self.vision_dim, initial_projection_dim, bias=True)
        self.fc2 = nn.Linear(initial_projection_dim, self.llm_dim, bias=True)
        self.fc3 = nn.Linear(self.llm_dim, self.llm_dim, bias=True)
        self.act_fn1 = nn.GELU()
        self.act_fn2 = nn.GELU()
        self.requires_grad_(False)
        self._init_native()

    def _linears(self):
        return self.fc1, self.fc2, self.fc3

    def forward(self, img_patches: torch.Tensor) -> torch.Tensor:
        return self.project(img_patches)
