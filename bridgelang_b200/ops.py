"""ops.py — torch-tensor front ends for the primitive C-ABI operators.

PyTorch is plumbing here (device memory + the current stream); every function hands raw pointers to
libbridgelang_b200.so and raises if the tensors are not CUDA tensors — there is no eager fallback.
"""

from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_QGELU, EPI_PATCH, EPI_RESIDUAL, Epilogue  # noqa: F401


_WORKSPACES: dict = {}


def shared_workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One grow-only scratch buffer per (device, stream): towers and the projector run back to back on a stream,
    so they can share it (their lifetimes never overlap); different streams get different buffers."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _WORKSPACES.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _stream() -> int:
    """Raw handle of the current stream of the CURRENT device; only valid inside `on_device(...)`."""
    return torch.cuda.current_stream().cuda_stream


@contextlib.contextmanager
def on_device(*ts: Optional[torch.Tensor]):
    """Device guard for a native call: every operand must live on ONE CUDA device, which becomes the current device
    for the duration of the call — kernels, TMA descriptors, per-device one-time setup and the stream handed to the
    library (`_stream()`) then all belong to the tensors' GPU, whatever device the caller had selected."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("bridgelang_b200 operators run on CUDA tensors only (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"operands on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("native call without a CUDA operand")
    with torch.cuda.device(dev):
        yield dev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bridgelang_b200 operators run on CUDA tensors only (no CPU fallback)")


def gemm(a: torch.Tensor, w: torch.Tensor, mode: int, *, bias=None, gamma=None, resid=None, out=None,
         out_col_off: int = 0, pos=None, tok_in: int = 0, tok_out: int = 0, tok_shift: int = 0,
         ln_stats=None, ln_colsum=None, ln_eps: float = 1e-6, stats_out=None, xb_out=None, shift_in=None,
         shift_out=None) -> Optional[torch.Tensor]:
    """C = a[M,K] @ w[N,K]^T with the fused epilogue `mode` (see include/bridgelang_b200.h).

    ln_stats [parts, M, 2] + ln_colsum [N]: LayerNorm folded into a BIAS / BIAS_GELU GEMM (a = bf16 copy of the
    un-normalised rows, w = W·diag(ln_w), bias = b + W·ln_b).  stats_out [gemm_stats_parts(N), M, 2] / xb_out [M, N]:
    what a RESIDUAL GEMM emits for such a consumer.  Rolling per-row shift: a RESIDUAL GEMM with shift_in [M] emits
    xb_out / stats_out of x_new − shift_in; a folded consumer with shift_in (+ shift_out [M]) writes the next
    producer's shift, shift_out = shift_in + mean(x − shift_in)."""
    _need_cuda(a, w, bias, gamma, resid, out, pos, ln_stats, ln_colsum, stats_out, xb_out, shift_in, shift_out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if mode in (EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_QGELU) and out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    e = Epilogue()
    e.bias, e.gamma, e.resid, e.pos = _ptr(bias), _ptr(gamma), _ptr(resid), _ptr(pos)
    e.ld_resid = resid.stride(0) if resid is not None else 0
    e.out = _ptr(out)
    e.ld_out = out.stride(0) if out is not None else 0
    e.out_col_off = out_col_off
    e.tok_in, e.tok_out, e.tok_shift = tok_in, tok_out, tok_shift
    if ln_stats is not None:
        assert ln_stats.dtype == torch.float32 and ln_stats.is_contiguous() and ln_stats.shape[1] == M
        e.ln_stats, e.ln_colsum, e.ln_parts, e.ln_eps = ln_stats.data_ptr(), _ptr(ln_colsum), ln_stats.shape[0], ln_eps
    if stats_out is not None:
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous()
        assert tuple(stats_out.shape) == (gemm_stats_parts(N), M, 2)
        e.stats_out = stats_out.data_ptr()
    if xb_out is not None:
        assert xb_out.dtype == torch.bfloat16 and xb_out.stride(1) == 1 and tuple(xb_out.shape) == (M, N)
        e.xb_out, e.ld_xb = xb_out.data_ptr(), xb_out.stride(0)
    if shift_in is not None:
        assert shift_in.dtype == torch.float32 and shift_in.numel() == M and shift_in.is_contiguous()
        e.shift_in = shift_in.data_ptr()
    if shift_out is not None:
        assert shift_out.dtype == torch.float32 and shift_out.numel() == M and shift_out.is_contiguous()
        e.shift_out = shift_out.data_ptr()
    lib = _lib.load()
    with on_device(a, w, bias, gamma, resid, out, pos, ln_stats, ln_colsum, stats_out, xb_out, shift_in, shift_out):
        _lib.check(lib.blb_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, mode, C.byref(e),
                                     _stream()), "gemm")
    return out


def gemm_stats_parts(n: int) -> int:
    """(sum, sumsq) pairs per row that an EPI_RESIDUAL GEMM with N = n writes into `stats_out`."""
    return int(_lib.load().blb_gemm_stats_parts(int(n)))


def rowstats_cast(x: torch.Tensor, parts: int, with_shift: bool = False):
    """fp32 rows → (bf16 copy, stats [parts, rows, 2] with the full-row (sum, sumsq) in part 0).  with_shift: both are
    those of x − c with c = the row mean, and the third return value is c [rows] (the primer of the rolling shift)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((parts, x.shape[0], 2), dtype=torch.float32, device=x.device)
    shift = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device) if with_shift else None
    with on_device(x):
        _lib.check(_lib.load().blb_rowstats_cast(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0),
                                                 stats.data_ptr(), parts, x.shape[0], x.shape[1], _ptr(shift),
                                                 _stream()), "rowstats_cast")
    return (y, stats, shift) if with_shift else (y, stats)


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    _need_cuda(x, weight, bias)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with on_device(x, weight, bias):
        _lib.check(_lib.load().blb_layernorm(x.data_ptr(), x.stride(0), weight.data_ptr(), bias.data_ptr(),
                                             y.data_ptr(), y.stride(0), x.shape[0], x.shape[1], eps, _stream()),
                   "layernorm")
    return y


def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int) -> torch.Tensor:
    """qkv: bf16 [batch*tokens, 3*heads*head_dim] (timm packing) → bf16 [batch*tokens, heads*head_dim]."""
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (batch * tokens, 3 * heads * head_dim)
    out = torch.empty((batch * tokens, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    with on_device(qkv):
        _lib.check(_lib.load().blb_attention(qkv.data_ptr(), out.data_ptr(), batch, tokens, heads, head_dim,
                                             _stream()), "attention")
    return out


def im2col_patch14(pixels: torch.Tensor, ldk: int = 592) -> torch.Tensor:
    _need_cuda(pixels)
    assert pixels.dtype == torch.bfloat16 and pixels.is_contiguous() and pixels.shape[1:] == (3, 224, 224)
    B = pixels.shape[0]
    cols = torch.empty((B * 256, ldk), dtype=torch.bfloat16, device=pixels.device)
    with on_device(pixels):
        _lib.check(_lib.load().blb_im2col_patch14(pixels.data_ptr(), cols.data_ptr(), B, ldk, _stream()), "im2col")
    return cols


_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def argmax(logits: torch.Tensor) -> torch.Tensor:
    """torch.argmax(logits, dim=-1) for a 2-D [rows, vocab] tensor, on the device, int64."""
    _need_cuda(logits)
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in _DTYPES
    ids = torch.empty((logits.shape[0],), dtype=torch.int64, device=logits.device)
    with on_device(logits):
        _lib.check(_lib.load().blb_argmax(logits.data_ptr(), _DTYPES[logits.dtype], logits.shape[0], logits.shape[1],
                                          logits.stride(0), ids.data_ptr(), _stream()), "argmax")
    return ids


def argmax_window(logits: torch.Tensor, begin: int, end: int) -> torch.Tensor:
    """torch.argmax(logits[:, begin:end], -1) + begin — the explicit windowed mode (e.g. the 256 action bins
    [31744, 32000)).  NOT the reference's greedy step, which arg-maxes the full row (`argmax`)."""
    _need_cuda(logits)
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in _DTYPES
    ids = torch.empty((logits.shape[0],), dtype=torch.int64, device=logits.device)
    with on_device(logits):
        _lib.check(_lib.load().blb_argmax_window(logits.data_ptr(), _DTYPES[logits.dtype], logits.shape[0],
                                                 logits.shape[1], logits.stride(0), int(begin), int(end),
                                                 ids.data_ptr(), _stream()), "argmax_window")
    return ids


class DecodeTables:
    """Device copies of bin_centers and the q01/q99/mask statistics of one dataset."""

    def __init__(self, bin_centers: np.ndarray, q01=None, q99=None, mask=None, device="cuda") -> None:
        self.bin_centers = torch.from_numpy(np.ascontiguousarray(bin_centers, dtype=np.float64)).to(device)
        self.q01 = None if q01 is None else torch.tensor(np.asarray(q01, dtype=np.float64), device=device)
        self.q99 = None if q99 is None else torch.tensor(np.asarray(q99, dtype=np.float64), device=device)
        self.mask = None if mask is None else torch.tensor(np.asarray(mask, dtype=bool), device=device).to(torch.uint8)
        self.action_dim = 0 if q01 is None else int(len(q01))


def detokenize_unnormalize(ids: torch.Tensor, vocab_size: int, tables: DecodeTables):
    """ids int64 [n] → (normalized float64 [n], actions float64 [n]); bit-exact vs the NumPy reference."""
    _need_cuda(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    n = ids.numel()
    norm = torch.empty((n,), dtype=torch.float64, device=ids.device)
    act = torch.empty((n,), dtype=torch.float64, device=ids.device)
    with on_device(ids, tables.bin_centers, tables.q01, tables.q99, tables.mask):
        _lib.check(_lib.load().blb_detokenize_unnormalize(
            ids.data_ptr(), n, vocab_size, tables.bin_centers.data_ptr(), tables.bin_centers.numel(),
            tables.action_dim, _ptr(tables.q01), _ptr(tables.q99), _ptr(tables.mask), norm.data_ptr(), act.data_ptr(),
            _stream()), "detokenize_unnormalize")
    return norm, act


def argmax_detokenize_unnormalize(logits: torch.Tensor, vocab_size: int, tables: DecodeTables):
    """One launch: full-row argmax → bin centre → un-normalize.  Returns (ids, normalized, actions)."""
    _need_cuda(logits)
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in _DTYPES
    rows = logits.shape[0]
    ids = torch.empty((rows,), dtype=torch.int64, device=logits.device)
    norm = torch.empty((rows,), dtype=torch.float64, device=logits.device)
    act = torch.empty((rows,), dtype=torch.float64, device=logits.device)
    with on_device(logits, tables.bin_centers, tables.q01, tables.q99, tables.mask):
        _lib.check(_lib.load().blb_argmax_detokenize_unnormalize(
            logits.data_ptr(), _DTYPES[logits.dtype], rows, logits.shape[1], logits.stride(0), vocab_size,
            tables.bin_centers.data_ptr(), tables.bin_centers.numel(), tables.action_dim, _ptr(tables.q01),
            _ptr(tables.q99), _ptr(tables.mask), ids.data_ptr(), norm.data_ptr(), act.data_ptr(), _stream()),
            "argmax_detokenize_unnormalize")
    return ids, norm, act


def argmax_window_detokenize_unnormalize(logits: torch.Tensor, begin: int, end: int, vocab_size: int,
                                         tables: DecodeTables):
    """Windowed argmax (one warp per row, shuffle reduction) → bin centre → un-normalize in one launch."""
    _need_cuda(logits)
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in _DTYPES
    rows = logits.shape[0]
    ids = torch.empty((rows,), dtype=torch.int64, device=logits.device)
    norm = torch.empty((rows,), dtype=torch.float64, device=logits.device)
    act = torch.empty((rows,), dtype=torch.float64, device=logits.device)
    with on_device(logits, tables.bin_centers, tables.q01, tables.q99, tables.mask):
        _lib.check(_lib.load().blb_argmax_window_detokenize_unnormalize(
            logits.data_ptr(), _DTYPES[logits.dtype], rows, logits.shape[1], logits.stride(0), int(begin), int(end),
            vocab_size, tables.bin_centers.data_ptr(), tables.bin_centers.numel(), tables.action_dim,
            _ptr(tables.q01), _ptr(tables.q99), _ptr(tables.mask), ids.data_ptr(), norm.data_ptr(), act.data_ptr(),
            _stream()), "argmax_window_detokenize_unnormalize")
    return ids, norm, act


def preprocess_u8(frames: torch.Tensor, lut: torch.Tensor):
    """uint8 [B,224,224,3] CUDA frames → (dino, siglip) bf16 [B,3,224,224]; lut bf16 [2,3,256] (see vision.py)."""
    _need_cuda(frames, lut)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and tuple(frames.shape[1:]) == (224, 224, 3)
    assert lut.dtype == torch.bfloat16 and lut.is_contiguous() and tuple(lut.shape) == (2, 3, 256)
    B = frames.shape[0]
    dino = torch.empty((B, 3, 224, 224), dtype=torch.bfloat16, device=frames.device)
    siglip = torch.empty_like(dino)
    with on_device(frames, lut):
        _lib.check(_lib.load().blb_preprocess_u8(frames.data_ptr(), B, lut.data_ptr(), dino.data_ptr(),
                                                 siglip.data_ptr(), _stream()), "preprocess_u8")
    return dino, siglip


_RESIZE_TABLES: dict = {}


def _resize_tables(in_size: int, out_size: int, interpolation: str, device: torch.device):
    from .resize import resample_coeffs
    key = (in_size, out_size, interpolation, device)
    t = _RESIZE_TABLES.get(key)
    if t is None:
        kk, bounds, ksize = resample_coeffs(in_size, out_size, interpolation)
        t = (torch.from_numpy(kk).to(device), torch.from_numpy(bounds).to(device), ksize)
        _RESIZE_TABLES[key] = t
    return t


def resize_u8(frames: torch.Tensor, out_hw, interpolation: str = "bicubic") -> torch.Tensor:
    """uint8 HWC frames [B,H,W,3] on the GPU → [B,Hd,Wd,3], bit-exact with PIL.Image.resize((Wd,Hd), BICUBIC) — the
    antialiased Resize of the reference's image transform (Pillow's fixed-point two-pass resampler on the device)."""
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_contiguous()
    B, Hs, Ws, _ = frames.shape
    Hd, Wd = int(out_hw[0]), int(out_hw[1])
    out = torch.empty((B, Hd, Wd, 3), dtype=torch.uint8, device=frames.device)
    if B == 0:
        return out
    kx, bx, ksx = _resize_tables(Ws, Wd, interpolation, frames.device) if Wd != Ws else (None, None, 0)
    ky, by, ksy = _resize_tables(Hs, Hd, interpolation, frames.device) if Hd != Hs else (None, None, 0)
    tmp = torch.empty((B, Hs, Wd, 3), dtype=torch.uint8, device=frames.device) if (Wd != Ws and Hd != Hs) else None
    with on_device(frames):
        _lib.check(_lib.load().blb_resize_u8(frames.data_ptr(), B, Hs, Ws, out.data_ptr(), Hd, Wd, _ptr(kx), _ptr(bx), ksx,
                                             _ptr(ky), _ptr(by), ksy, _ptr(tmp), _stream()), "resize_u8")
    return out


def u8_to_patches(frames: torch.Tensor, ldk: int = 592) -> torch.Tensor:
    """uint8 HWC frames [B,S,S,3] (S a multiple of 14) → bf16 [B*(S/14)², ldk], k = kh*42 + kw*3 + c, exact 0..255."""
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_contiguous()
    B, S = frames.shape[0], frames.shape[1]
    grid = S // 14
    cols = torch.empty((B * grid * grid, ldk), dtype=torch.bfloat16, device=frames.device)
    with on_device(frames):
        _lib.check(_lib.load().blb_u8_to_patches(frames.data_ptr(), cols.data_ptr(), B, ldk, grid, S, _stream()),
                   "u8_to_patches")
    return cols


def encode_actions(actions: torch.Tensor, bins: torch.Tensor, min_action: float, max_action: float,
                   vocab_size: int) -> torch.Tensor:
    """clip → np.digitize(bins) → vocab_size − index, on the device; actions float32/float64, any shape."""
    _need_cuda(actions, bins)
    assert actions.dtype in (torch.float32, torch.float64) and bins.dtype == torch.float64 and bins.is_contiguous()
    a = actions.contiguous()
    ids = torch.empty(a.shape, dtype=torch.int64, device=a.device)
    if a.numel() == 0:
        return ids
    dt = _lib.DTYPE_F32 if a.dtype == torch.float32 else _lib.DTYPE_F64
    with on_device(a, bins):
        _lib.check(_lib.load().blb_encode_actions(a.data_ptr(), dt, a.numel(), bins.data_ptr(), bins.numel(),
                                                  float(min_action), float(max_action), int(vocab_size),
                                                  ids.data_ptr(), _stream()), "encode_actions")
    return ids


def action_token_metrics(logits: torch.Tensor, labels: torch.Tensor, num_patches: int, action_token_begin_idx: int,
                         vocab_size: int, tables: "DecodeTables"):
    """base_strategy.py:314-329 on the device.  Returns dict(preds [B, L-1] (−1 where unmasked), counts int64 [2] =
    (#correct, #mask), l1_sum float64 [1]); accuracy = counts[0]/counts[1], l1 = l1_sum/counts[1] (no host sync here)."""
    _need_cuda(logits, labels)
    assert logits.dim() == 3 and logits.stride(2) == 1 and logits.dtype in _DTYPES
    assert labels.dim() == 2 and labels.dtype == torch.int64 and labels.stride(1) == 1
    B, S, V = logits.shape
    n_pos = S - 1 - num_patches
    assert n_pos > 0 and labels.shape[0] == B and labels.shape[1] >= n_pos + 1
    preds = torch.empty((B, n_pos), dtype=torch.int64, device=logits.device)
    absdiff = torch.empty((B, n_pos), dtype=torch.float64, device=logits.device)
    counts = torch.empty((2,), dtype=torch.int64, device=logits.device)
    l1_sum = torch.empty((1,), dtype=torch.float64, device=logits.device)
    with on_device(logits, labels, tables.bin_centers):
        _lib.check(_lib.load().blb_action_token_metrics(
            logits.data_ptr(), _DTYPES[logits.dtype], B, S, V, logits.stride(1), logits.stride(0), num_patches,
            labels.data_ptr(), labels.stride(0), int(action_token_begin_idx), int(vocab_size),
            tables.bin_centers.data_ptr(), tables.bin_centers.numel(), preds.data_ptr(), absdiff.data_ptr(),
            counts.data_ptr(), l1_sum.data_ptr(), _stream()), "action_token_metrics")
    return {"preds": preds, "absdiff": absdiff, "counts": counts, "l1_sum": l1_sum}


def launch_count() -> int:
    return int(_lib.load().blb_launch_count())


def set_gemm_cta_group(ctas: int) -> None:
    _lib.load().blb_set_gemm_cta_group(int(ctas))


TIMING_CATEGORIES = ("gemm", "attention", "layernorm", "other")


def timing_enable(on: bool) -> None:
    _lib.load().blb_timing_enable(1 if on else 0)


def timing_reset() -> None:
    _lib.load().blb_timing_reset()


def timing_collect() -> dict:
    """{category: {"ms", "work", "launches"}} summed over every launch recorded since the last reset.
    Call after torch.cuda.synchronize()."""
    lib = _lib.load()
    out = {}
    for i, name in enumerate(TIMING_CATEGORIES):
        ms, work, n = C.c_double(), C.c_double(), C.c_longlong()
        _lib.check(lib.blb_timing_collect(i, C.byref(ms), C.byref(work), C.byref(n)), "timing_collect")
        out[name] = {"ms": ms.value, "work": work.value, "launches": n.value}
    return out


def timing_records(max_records: int = 4096) -> list:
    """Every instrumented launch since the last reset, in launch order: dicts {cat, ms, work} and, for GEMMs,
    {mode, ln_folded, emits_stats, N, K} decoded from the tag.  Call after torch.cuda.synchronize()."""
    cat = (C.c_int * max_records)()
    tag = (C.c_longlong * max_records)()
    ms = (C.c_double * max_records)()
    work = (C.c_double * max_records)()
    n = _lib.load().blb_timing_records(max_records, cat, tag, ms, work)
    if n < 0:
        raise RuntimeError(f"blb_timing_records failed: {n}")
    out = []
    for i in range(n):
        r = {"cat": TIMING_CATEGORIES[cat[i]], "ms": ms[i], "work": work[i]}
        if cat[i] == 0:
            t = tag[i]
            r.update(mode=t >> 44, ln_folded=(t >> 43) & 1, emits_stats=(t >> 42) & 1, N=(t >> 20) & 0xFFFFF,
                     K=t & 0xFFFFF)
        out.append(r)
    return out
