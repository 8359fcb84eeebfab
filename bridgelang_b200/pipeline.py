"""pipeline.py — the fused featurize+project call and its data-parallel sharding.

`VisualPrefixEncoder` is what `PrismaticVLM.forward` does at prismatic/models/vlms/prismatic.py:367-375
(`vision_backbone(pixel_values)` then `projector(...)`) as ONE C-ABI call
(`blb_fused_featurize_project_forward`): DINOv2 tower → SigLIP tower → 3-GEMM projector, sharing one workspace.

Multi-GPU: every image is independent, so the path shards by image batch — one process per GPU,
weights replicated, contiguous image slices per rank and NO collective inside the path.  The only exchange is the
optional re-assembly of projected prefixes when the LLM runs on fewer ranks: an NCCL all-gather over NVLink
(`gather_prefixes`; gloo on CPU for the host-logic tests).
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, ops
from .config import NUM_PATCHES
from .projector import FusedMLPProjector, PrismaticProjector
from .vision import DinoSigLIPViTBackbone, PrismaticVisionBackbone, _as_pixels, _FusedBackbone


class VisualPrefixEncoder(nn.Module):
    """pixel_values {"dino","siglip"} (or HF-packed [B,6,224,224]) → projected prefix [B, 256, llm_dim] bf16."""

    def __init__(self, vision_backbone: nn.Module, projector: nn.Module) -> None:
        super().__init__()
        self.vision_backbone = vision_backbone
        self.projector = projector
        if isinstance(vision_backbone, _FusedBackbone):          # DinoSigLIP (224 / 384 px) and DinoCLIP (336 px)
            self._towers = vision_backbone._towers()
            self._keys = vision_backbone.KEYS
        elif isinstance(vision_backbone, PrismaticVisionBackbone) and vision_backbone.use_fused_vision_backbone:
            self._towers = (vision_backbone.featurizer, vision_backbone.fused_featurizer)
        else:
            raise ValueError("VisualPrefixEncoder needs a fused DINOv2+SigLIP backbone")
        if not isinstance(projector, (FusedMLPProjector, PrismaticProjector)):
            raise ValueError("VisualPrefixEncoder needs a FusedMLPProjector / PrismaticProjector")

    def _split(self, pixel_values) -> Tuple[torch.Tensor, torch.Tensor]:
        if isinstance(pixel_values, dict):
            k0, k1 = getattr(self, "_keys", ("dino", "siglip"))
            return pixel_values[k0], pixel_values[k1]
        img, img_fused = torch.split(pixel_values, [3, 3], dim=1)   # modeling_prismatic.py:120
        return img, img_fused

    @torch.no_grad()
    def forward(self, pixel_values, return_features: bool = False):
        dino_px, siglip_px = self._split(pixel_values)
        cfg0, cfg1 = self._towers[0].cfg, self._towers[1].cfg
        dino_px, siglip_px = _as_pixels(dino_px, cfg0.img_size), _as_pixels(siglip_px, cfg1.img_size)
        B, dev = dino_px.shape[0], dino_px.device
        lib = _lib.load()
        dino, siglip = self._towers[0].packed(), self._towers[1].packed()
        proj = self.projector.packed()
        fused_dim = dino.struct.dim + siglip.struct.dim
        feats = torch.empty((B, cfg0.num_patches, fused_dim), dtype=torch.bfloat16, device=dev)
        out = torch.empty((B, cfg0.num_patches, proj.out_dim), dtype=torch.bfloat16, device=dev)
        if B == 0:
            return (out, feats) if return_features else out
        need = lib.blb_fused_workspace_bytes(C.byref(dino.struct), C.byref(siglip.struct), C.byref(proj), B)
        with ops.on_device(dino_px, siglip_px, self._towers[0].pos_embed, self._towers[1].pos_embed, feats, out):
            ws = ops.shared_workspace(dev, need)
            _lib.check(lib.blb_fused_featurize_project_forward(
                C.byref(dino.struct), C.byref(siglip.struct), C.byref(proj), dino_px.data_ptr(), siglip_px.data_ptr(),
                B, feats.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream().cuda_stream), "fused_featurize_project_forward")
        return (out, feats) if return_features else out

    @torch.no_grad()
    def forward_uint8(self, frames: torch.Tensor, return_features: bool = False, folded: bool = True):
        """Already-resized uint8 frames [B,224,224,3] on the GPU → projected prefix.

        folded=True (default): ToTensor and each tower's Normalize live inside that tower's patch-embed weights, and —
        stride == kernel making im2col a pure permutation — the frame is written ONCE in patch-major order as exact bf16
        integers: that matrix is the TMA-loaded A operand of BOTH towers' patch-embed GEMMs (one C-ABI call,
        `blb_fused_featurize_project_forward_u8`).  No normalized frames, no per-tower im2col pass.
        folded=False: the round-1 route — a LUT kernel writes both normalized bf16 frames (bit-identical to the host
        transform), then exactly `forward`."""
        if not folded:
            if not isinstance(self.vision_backbone, DinoSigLIPViTBackbone):
                raise ValueError("forward_uint8(folded=False) needs the native DinoSigLIPViTBackbone (dict pixel_values)")
            return self.forward(self.vision_backbone.preprocess_uint8(frames), return_features=return_features)
        from .vision import _as_frames
        cfg = self._towers[0].cfg
        fr = _as_frames(frames, cfg.img_size)
        B, dev = fr.shape[0], fr.device
        lib = _lib.load()
        dino, siglip = self._towers[0].packed(), self._towers[1].packed()
        proj = self.projector.packed()
        fused_dim = dino.struct.dim + siglip.struct.dim
        feats = torch.empty((B, cfg.num_patches, fused_dim), dtype=torch.bfloat16, device=dev)
        out = torch.empty((B, cfg.num_patches, proj.out_dim), dtype=torch.bfloat16, device=dev)
        if B == 0:
            return (out, feats) if return_features else out
        fused = lib.blb_fused_workspace_bytes(C.byref(dino.struct), C.byref(siglip.struct), C.byref(proj), B)
        need = (fused + 255) // 256 * 256 + lib.blb_patch_matrix_bytes(C.byref(dino.struct), B)
        with ops.on_device(fr, self._towers[0].pos_embed, self._towers[1].pos_embed, feats, out):
            ws = ops.shared_workspace(dev, need)
            _lib.check(lib.blb_fused_featurize_project_forward_u8(
                C.byref(dino.struct), C.byref(siglip.struct), C.byref(proj), fr.data_ptr(), B, feats.data_ptr(),
                out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                "fused_featurize_project_forward_u8")
        return (out, feats) if return_features else out

    def stream(self, host_batches, uint8: bool = False, to_host: bool = False):
        """Serving loop with double-buffered input staging: yields the projected prefix of every batch in
        `host_batches` (an iterable of pinned host `pixel_values` dicts / HF-packed tensors, or uint8 frame tensors
        with `uint8=True`), copying batch i+1 host→device on a side stream while batch i is being encoded.  Same
        results as calling `forward` per batch; the H2D copy just leaves the critical path.

        `to_host=True` also brings every result back: the prefix is copied device→host into one of two pinned buffers
        on a third stream (under the next batch's encode) and the generator yields `(host_tensor, done_event)` —
        synchronize the event before reading; the buffer is reused two batches later."""
        dev = next(self.parameters()).device
        compute = torch.cuda.current_stream(dev)
        copier = torch.cuda.Stream(dev)
        slots = [None, None]                      # device staging buffers, reused
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        drained = [torch.cuda.Event(), torch.cuda.Event()]
        used = [False, False]
        drainer = torch.cuda.Stream(dev) if to_host else None
        host_out = [None, None]
        out_done = [torch.cuda.Event(), torch.cuda.Event()]
        keep_alive = [None, None]                 # device results whose D2H copy may still be running

        def alloc_like(v):
            # Allocate the staging buffer ON the copier stream: the caching allocator then never hands out a block
            # that kernels already queued on the compute stream (e.g. the consumer's work on an earlier result) may
            # still be reading, and record_stream tells it that the compute stream reads the buffer as well.
            with torch.cuda.stream(copier):
                t = torch.empty_like(v, device=dev)
            t.record_stream(compute)
            return t

        def upload(i, batch):
            # staging buffers mirror the host tensors' strides (empty_like): a layout mismatch would turn the async
            # H2D memcpy into a host-side re-layout plus a synchronous staged copy (measured: +15 ms per 154 MB batch)
            if isinstance(batch, dict):
                if slots[i] is None or any(slots[i][k].shape != v.shape or slots[i][k].stride() != v.stride()
                                           for k, v in batch.items()):
                    slots[i] = {k: alloc_like(v) for k, v in batch.items()}
            elif slots[i] is None or slots[i].shape != batch.shape or slots[i].stride() != batch.stride():
                slots[i] = alloc_like(batch)
            with torch.cuda.stream(copier):
                if used[i]:
                    copier.wait_event(drained[i])  # the encoder finished reading this slot
                if isinstance(batch, dict):
                    for k, v in batch.items():
                        slots[i][k].copy_(v, non_blocking=True)
                else:
                    slots[i].copy_(batch, non_blocking=True)
                ready[i].record(copier)

        def download(i, out):
            if host_out[i] is None or host_out[i].shape != out.shape:
                host_out[i] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            done = torch.cuda.Event()
            done.record(compute)                   # the projector's last kernel
            with torch.cuda.stream(drainer):
                drainer.wait_event(done)
                host_out[i].copy_(out, non_blocking=True)
                out_done[i] = torch.cuda.Event()
                out_done[i].record(drainer)
            out.record_stream(drainer)
            keep_alive[i] = out
            return host_out[i], out_done[i]

        it = iter(host_batches)
        try:
            upload(0, next(it))
        except StopIteration:
            return
        i = 0
        while True:
            nxt = next(it, None)
            if nxt is not None:
                upload(1 - i, nxt)                 # overlaps with the encode below
            compute.wait_event(ready[i])
            out = self.forward_uint8(slots[i]) if uint8 else self.forward(slots[i])
            drained[i].record(compute)
            used[i] = True
            if to_host:
                if host_out[i] is not None:
                    out_done[i].synchronize()      # the consumer had a full batch of time to read this buffer
                yield download(i, out)
            else:
                yield out
            if nxt is None:
                return
            i = 1 - i


# ----------------------------------------------------------------------------------------------------------
# data-parallel sharding
# ----------------------------------------------------------------------------------------------------------
def shard_bounds(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous image slice [lo, hi) of rank `rank`; the first (global_batch % world_size) ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size) or global_batch < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pixel_values(pixel_values: Dict[str, torch.Tensor], rank: int, world_size: int) -> Dict[str, torch.Tensor]:
    n = next(iter(pixel_values.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world_size)
    return {k: v[lo:hi] for k, v in pixel_values.items()}


def _gather_layout(global_batch: int, world: int):
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    return sizes, max(hi - lo for lo, hi in sizes)


def _gather_into(local: torch.Tensor, global_batch: int, group) -> torch.Tensor:
    """Enqueue the all-gather on the current stream; uneven shards are staged into one max-shard-sized send buffer
    (a single allocation + one copy, the pad rows are never read by the caller)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes, max_n = _gather_layout(global_batch, world)
    lo, hi = sizes[rank]
    assert local.shape[0] == hi - lo, "local shard does not match shard_bounds()"
    send = local.contiguous()
    if send.shape[0] < max_n:
        staged = torch.empty((max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
        staged[:send.shape[0]].copy_(send)
        send = staged
    gathered = torch.empty((world * max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, send, group=group)
    if all(h - l == max_n for l, h in sizes):
        return gathered
    parts: List[torch.Tensor] = [gathered[r * max_n: r * max_n + (h - l)] for r, (l, h) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def gather_prefixes(local: torch.Tensor, global_batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank projected prefixes [B_local, 256, llm_dim] into [global_batch, 256, llm_dim] on
    every rank (NCCL over NVLink on GPUs), in line on the current stream."""
    return _gather_into(local, global_batch, group)


class PrefixGatherer:
    """The same all-gather, taken off the critical path: `launch(local)` enqueues it on a communication stream (which
    waits for the producer of `local`), `wait(handle)` makes the current stream wait for it and returns the gathered
    tensor.  A serving loop calls launch() for step i, encodes step i+1, then wait()s — the 2 MiB per image that cross
    NVLink (3.5 GiB received per rank at 8 x 256 images) move under the next step's GEMMs.

    mode="p2p" (default on CUDA): every rank PUSHES its shard straight into every peer's gather buffer through NVLink
    peer memory (`torch.distributed._symmetric_memory`: one symmetric allocation, plain device-to-device copies, i.e.
    the copy engines) — no SM, no shared-memory and no register footprint next to the towers' persistent kernels, where
    an NCCL all-gather kernel costs 2.5-4.6 % of the 8-GPU step even when overlapped (profiles/r02_scaling.md).  Two
    cross-rank barriers bracket the pushes: nobody overwrites a buffer a peer may still be reading, nobody reads before
    every peer has written.  mode="nccl": `all_gather_into_tensor` on the communication stream.  CPU tensors (gloo, the
    host-logic tests) run in line."""

    def __init__(self, global_batch: int, group: Optional[dist.ProcessGroup] = None, mode: str = "p2p") -> None:
        if mode not in ("p2p", "nccl"):
            raise ValueError("mode must be 'p2p' or 'nccl'")
        self.global_batch, self.group, self.mode = global_batch, group, mode
        self._stream: Optional[torch.cuda.Stream] = None
        self._symm = None          # (buffers[2], handles[2], peer views[2][world]) once allocated
        self._slot = 0

    # -- p2p transport ---------------------------------------------------------------------------------
    def _symm_buffers(self, local: torch.Tensor, world: int, max_n: int):
        if self._symm is None:
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else dist.group.WORLD
            shape = (world * max_n, *local.shape[1:])
            bufs, hdls, views = [], [], []
            for _ in range(2):     # double-buffered: the consumer reads gather i while gather i+1 is being pushed
                b = symm.empty(shape, dtype=local.dtype, device=local.device)
                h = symm.rendezvous(b, grp.group_name)
                bufs.append(b)
                hdls.append(h)
                views.append([h.get_buffer(r, shape, local.dtype) for r in range(world)])
            self._symm = (bufs, hdls, views)
        return self._symm

    def _push(self, local: torch.Tensor):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        sizes, max_n = _gather_layout(self.global_batch, world)
        lo, hi = sizes[rank]
        assert local.shape[0] == hi - lo, "local shard does not match shard_bounds()"
        bufs, hdls, views = self._symm_buffers(local, world, max_n)
        s = self._slot
        self._slot ^= 1
        hdls[s].barrier(channel=0)                       # every rank's consumer is done with slot s (launched 2 gathers ago)
        src = local.contiguous()
        for k in range(world):                           # start with the own copy, then walk the ring
            r = (rank + k) % world
            views[s][r][rank * max_n: rank * max_n + src.shape[0]].copy_(src, non_blocking=True)
        hdls[s].barrier(channel=1)                       # every peer's shard has landed in this rank's buffer
        out = bufs[s]
        if all(h - l == max_n for l, h in sizes):
            return out[: self.global_batch]
        return torch.cat([out[r * max_n: r * max_n + (h - l)] for r, (l, h) in enumerate(sizes)], dim=0)

    def launch(self, local: torch.Tensor):
        if not local.is_cuda:
            return (_gather_into(local, self.global_batch, self.group), None)
        if self._stream is None or self._stream.device != local.device:
            self._stream = torch.cuda.Stream(local.device)
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(local.device))   # also orders the consumer's earlier reads of the slot
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(produced)
            if self.mode == "p2p":
                out = self._push(local)
            else:
                out = _gather_into(local, self.global_batch, self.group)
            done = torch.cuda.Event()
            done.record(self._stream)
        local.record_stream(self._stream)
        return (out, done)

    def wait(self, handle) -> torch.Tensor:
        out, done = handle
        if done is not None:
            torch.cuda.current_stream(out.device).wait_event(done)
            if self.mode != "p2p":
                out.record_stream(torch.cuda.current_stream(out.device))
        return out
