"""step_breakdown.py — where the featurize+project step spends its time AND its power budget (development tool).

    python tools/step_breakdown.py [--batch 256] [--loops]

1. per-kernel-shape table of one step (CUDA events around every launch, `blb_timing_records`) for the LN-folded and
   the explicit-LayerNorm schedules, same process, same box;
2. NVML power / SM clock sampled every ~2 ms during un-instrumented steps of each schedule;
3. --loops: each dominant kernel shape alone in a 1.5 s loop → its SUSTAINED (power-capped) throughput, power and
   SM clock, next to torch.matmul (cuBLAS) on the same shape.
"""

from __future__ import annotations

import argparse
import statistics
import sys
import threading
import time
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import build_encoder  # noqa: E402
from bridgelang_b200 import ops  # noqa: E402
from bridgelang_b200.weights import normalize_frames, synthetic_frames  # noqa: E402


class Nvml:
    """power (W) and SM clock (MHz) every ~2 ms on a thread."""

    def __init__(self):
        import pynvml
        self.n = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.rows, self._stop, self._t = [], threading.Event(), None

    def start(self):
        self.rows, self._stop = [], threading.Event()

        def run():
            while not self._stop.is_set():
                try:
                    self.rows.append((self.n.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                                      self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM)))
                except Exception:
                    pass
                time.sleep(0.002)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join()
        if not self.rows:
            return {}
        p, c = [r[0] for r in self.rows], [r[1] for r in self.rows]
        return {"n": len(p), "power_mean": statistics.fmean(p), "power_max": max(p), "sm_mhz_mean": statistics.fmean(c),
                "sm_mhz_min": min(c), "sm_mhz_max": max(c)}


MODE = {0: "bias", 1: "gelu", 2: "resid", 3: "patch"}


def shape_table(enc, px, steps=2):
    ops.timing_enable(True)
    ops.timing_reset()
    torch.cuda.synchronize()
    for _ in range(steps):
        enc(px)
    torch.cuda.synchronize()
    recs = ops.timing_records(8192)
    ops.timing_enable(False)
    ops.timing_reset()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in recs:
        if r["cat"] == "gemm":
            key = f"gemm {MODE[r['mode']]:5s} N={r['N']:<5d} K={r['K']:<5d}" + (" +ln" if r["ln_folded"] else "") + \
                  (" +stats" if r["emits_stats"] else "")
        else:
            key = r["cat"]
        a = agg[key]
        a[0] += 1
        a[1] += r["ms"]
        a[2] += r["work"]
    total = sum(a[1] for a in agg.values()) / steps
    print(f"{'kernel':44s} {'n/step':>6s} {'ms/step':>8s} {'share':>6s} {'T(FLOP|B)/s':>11s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {a[0] // steps:6d} {a[1] / steps:8.3f} {100 * a[1] / steps / total:5.1f}% {a[2] / a[1] / 1e9:11.1f}")
    print(f"{'sum of launches':44s} {'':6s} {total:8.3f}")
    return total


def timed_steps(enc, px, steps=10, warmup=3, nv=None):
    for _ in range(warmup):
        enc(px)
    torch.cuda.synchronize()
    if nv:
        nv.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        enc(px)
    e.record()
    torch.cuda.synchronize()
    info = nv.stop() if nv else {}
    return s.elapsed_time(e) / steps, info


def sustained(name, fn, flops, nv, seconds=1.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    nv.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, n = time.perf_counter(), 0
    s.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    info = nv.stop()
    ms = s.elapsed_time(e) / n
    print(f"{name:40s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {info.get('power_mean', 0):6.0f} W  "
          f"{info.get('sm_mhz_mean', 0):6.0f} MHz")


def loops(nv, cublas=True):
    g = torch.Generator(device="cuda").manual_seed(0)
    M = 66816
    for (label, N, K, mode) in (("qkv  bias", 3072, 1024, 0), ("fc1  gelu", 4096, 1024, 1), ("fc2  resid", 1024, 4096, 2),
                                ("proj resid", 1024, 1024, 2)):
        a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g) * 0.1
        flops = 2.0 * M * N * K
        if mode == 2:
            resid = torch.zeros(M, N, device="cuda")
            gamma = torch.full((N,), 1e-3, device="cuda")
            sustained(f"ours  {label} N={N} K={K}", lambda: ops.gemm(a, w, 2, bias=bias, gamma=gamma, resid=resid), flops, nv)
            parts = ops.gemm_stats_parts(N)
            stats = torch.empty(parts, M, 2, device="cuda")
            xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            sustained(f"ours  {label} N={N} K={K} +stats", lambda: ops.gemm(a, w, 2, bias=bias, gamma=gamma, resid=resid,
                                                                          stats_out=stats, xb_out=xb), flops, nv)
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            sustained(f"ours  {label} N={N} K={K}", lambda: ops.gemm(a, w, mode, bias=bias, out=out), flops, nv)
            _, st = ops.rowstats_cast(a.float(), ops.gemm_stats_parts(K))
            cs = w.float().sum(dim=1)
            sustained(f"ours  {label} N={N} K={K} +ln", lambda: ops.gemm(a, w, mode, bias=bias, out=out, ln_stats=st,
                                                                       ln_colsum=cs), flops, nv)
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        wt = w.t()
        if cublas:
            sustained(f"cuBLAS plain     N={N} K={K}", lambda: torch.matmul(a, wt, out=out2), flops, nv)
        del a, w
    if not cublas:
        return
    a = torch.randn(8192, 8192, device="cuda", generator=g).bfloat16()
    b = torch.randn(8192, 8192, device="cuda", generator=g).bfloat16()
    o = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
    sustained("cuBLAS 8192^3", lambda: torch.matmul(a, b, out=o), 2.0 * 8192 ** 3, nv)
    qkv = torch.randn(256 * 261, 3072, device="cuda", generator=g).bfloat16()
    sustained("attention dino 261x64", lambda: ops.attention(qkv, 256, 261, 16, 64), 4.0 * 256 * 16 * 261 * 261 * 64, nv)
    qkv = torch.randn(256 * 256, 3456, device="cuda", generator=g).bfloat16()
    sustained("attention siglip 256x72", lambda: ops.attention(qkv, 256, 256, 16, 72), 4.0 * 256 * 16 * 256 * 256 * 72, nv)
    x = torch.randn(66816, 1024, device="cuda", generator=g)
    w1, b1 = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    sustained("layernorm 66816x1024 (GB/s)", lambda: ops.layernorm(x, w1, b1), 6.0 * 66816 * 1024, nv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--loops", action="store_true")
    ap.add_argument("--loops-only", action="store_true")
    ap.add_argument("--no-cublas", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    nv = Nvml()
    if args.loops_only:
        loops(nv, not args.no_cublas)
        return
    enc = build_encoder(dev)
    px = {k: v.to(torch.bfloat16).to(dev) for k, v in normalize_frames(synthetic_frames(args.batch, seed=1000)).items()}
    towers = (enc.vision_backbone.dino_featurizer, enc.vision_backbone.siglip_featurizer)
    for folded in (True, False, True):
        for t in towers:
            t.ln_folded = folded
            t.invalidate_packed()
        ms, info = timed_steps(enc, px, nv=nv)
        print(f"\n=== ln_folded={folded}: {ms:.2f} ms/step = {args.batch / ms * 1e3:.0f} img/s   nvml {info}")
        shape_table(enc, px)
    if args.loops:
        print("\n=== sustained loops (1.5 s each) ===")
        loops(nv)


if __name__ == "__main__":
    main()
