"""bringup.py — per-kernel smoke/parity/timing on a real B200 (development tool, not part of the test suite).

usage: python tools/bringup.py <case> [args]     (each case is run in its own process under `timeout`)
"""

from __future__ import annotations

import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bridgelang_b200 import ops  # noqa: E402


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def time_it(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def case_gemm_fold(ctas: int, M: int, N: int, K: int, mode: int):
    """the LN-folded variants as the towers run them: mode 0/1 consume (stats, colsum), mode 2 emits (stats, xb)."""
    ops.set_gemm_cta_group(ctas)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    if mode in (ops.EPI_BIAS, ops.EPI_BIAS_GELU):
        _, st = ops.rowstats_cast(a.float(), ops.gemm_stats_parts(K))
        cs = w.float().sum(dim=1)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, w, mode, bias=bias, out=out, ln_stats=st, ln_colsum=cs)
    else:
        gamma = torch.full((N,), 1e-3, device="cuda")
        resid = torch.zeros(M, N, device="cuda")
        stats = torch.empty(ops.gemm_stats_parts(N), M, 2, device="cuda")
        xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, w, 2, bias=bias, gamma=gamma, resid=resid, stats_out=stats, xb_out=xb)
    ms = time_it(fn)
    print(f"gemm_fold ctas={ctas} M={M} N={N} K={K} mode={mode}  time {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def case_gemm(ctas: int, M: int, N: int, K: int, mode: int, check: bool = True):
    ops.set_gemm_cta_group(ctas)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    resid0 = torch.randn(M, N, device="cuda", generator=g)
    if mode in (ops.EPI_BIAS, ops.EPI_BIAS_GELU):
        out = ops.gemm(a, w, mode, bias=bias)
        torch.cuda.synchronize()
        if check:
            ref = a.float() @ w.float().t() + bias
            if mode == ops.EPI_BIAS_GELU:
                ref = torch.nn.functional.gelu(ref)
            print(f"gemm ctas={ctas} M={M} N={N} K={K} mode={mode} relerr={relerr(out, ref):.3e}")
        fn = lambda: ops.gemm(a, w, mode, bias=bias, out=out)
    else:
        resid = resid0.clone()
        ops.gemm(a, w, ops.EPI_RESIDUAL, bias=bias, gamma=gamma, resid=resid)
        torch.cuda.synchronize()
        if check:
            ref = resid0 + gamma * (a.float() @ w.float().t() + bias)
            print(f"gemm ctas={ctas} M={M} N={N} K={K} mode=resid relerr={relerr(resid, ref):.3e}")
        fn = lambda: ops.gemm(a, w, ops.EPI_RESIDUAL, bias=bias, gamma=gamma, resid=resid)
    ms = time_it(fn)
    print(f"   time {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def case_ln():
    g = torch.Generator(device="cuda").manual_seed(0)
    for D in (1024, 1152):
        x = torch.randn(66816, D, device="cuda", generator=g) * 3 + 1
        w = torch.rand(D, device="cuda", generator=g) + 0.5
        b = torch.randn(D, device="cuda", generator=g) * 0.1
        y = ops.layernorm(x, w, b, 1e-6)
        ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)
        ms = time_it(lambda: ops.layernorm(x, w, b, 1e-6))
        print(f"layernorm D={D} relerr={relerr(y, ref):.3e} time {ms*1e3:.1f} us  {x.numel()*6/ms/1e6:.0f} GB/s")


def case_attn():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, T, H, hd) in ((4, 261, 16, 64), (4, 256, 16, 72), (256, 261, 16, 64), (256, 256, 16, 72)):
        D = H * hd
        qkv = (torch.randn(B * T, 3 * D, device="cuda", generator=g)).bfloat16()
        out = ops.attention(qkv, B, T, H, hd)
        torch.cuda.synchronize()
        if B <= 8:
            q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
            ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
            print(f"attention B={B} T={T} hd={hd} relerr={relerr(out, ref):.3e}")
        ms = time_it(lambda: ops.attention(qkv, B, T, H, hd))
        print(f"   time {ms:.3f} ms  {4.0*B*H*T*T*hd/ms/1e9:.1f} TFLOP/s")


def case_im2col():
    g = torch.Generator(device="cuda").manual_seed(0)
    px = torch.randn(3, 3, 224, 224, device="cuda", generator=g).bfloat16()
    cols = ops.im2col_patch14(px)
    ref = torch.nn.functional.unfold(px.float(), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    print("im2col exact:", bool((cols[:, :588].float() == ref).all()), "pad zero:", bool((cols[:, 588:] == 0).all()))


def case_argmax():
    g = torch.Generator(device="cuda").manual_seed(7)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        x = torch.randn(7, 32064, device="cuda", generator=g).to(dt)
        x[1, 100] = x[1].max() ; x[1, 50] = x[1, 100]
        ids = ops.argmax(x)
        print("argmax", dt, bool((ids == torch.argmax(x, dim=-1)).all()), ids.tolist())


if __name__ == "__main__":
    case = sys.argv[1]
    if case == "gemm_fold":
        case_gemm_fold(*map(int, sys.argv[2:7]))
    elif case == "gemm":
        ctas, M, N, K, mode = map(int, sys.argv[2:7])
        case_gemm(ctas, M, N, K, mode, check=(M * N <= 70000 * 4400))
    else:
        globals()["case_" + case]()
