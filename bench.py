#!/usr/bin/env python
"""bench.py — images/s of the prism-dinosiglip-224px featurize+project path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--gather] [--impl reference]

A step = one featurize+project forward over one batch of B=256 synthetic 224x224 frames per GPU (BASELINE config
"prism-dinosiglip-224px featurizer + projector, bf16, batch 256 synthetic 224px frames, 1xB200"); N>1 shards a
global batch of 256*N images by image (weak scaling, no collective in the path; --gather adds the optional NCCL
all-gather of the projected prefixes).  One JSON line is printed by rank 0:

  value      whole-job images/s with the normalized bf16 frames already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API (VisualPrefixEncoder) from PINNED HOST frames, H2D copy and a D2H
             read of a per-image digest inside the timed region
  roofline   the dominant kernel family (tcgen05 GEMM): algorithmic FLOPs / device time of those launches, measured
             live with CUDA events around every launch of a separate instrumented pass, vs MEASURED_PEAKS.json
  cpu_baseline  the fp32 oracle restatement of the reference path (all 24+27 blocks, as timm executes) on the box's
             host cores, rank 0, N=1 only
--impl reference times that same CPU restatement (the reference package itself cannot be imported offline: timm,
draccus, tensorflow... are absent) on all host threads.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "images/sec DinoSigLIP-224px featurize+project bf16 b=256"
UNIT = "images/s"
WORKLOAD = "prism-dinosiglip-224px featurizer + FusedMLPProjector, bf16, batch 256 synthetic 224px frames per GPU"


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r01_ncu_full_c_kernels.md), keyed like ops.timing_records: (mode, N, K, ln_folded, emits_stats)
NCU_DRAM_BYTES_PER_LAUNCH = {
    (0, 3072, 1024, 1, 0): 147.46e6 + 362.11e6,     # DINOv2 qkv   (algorithmic: A 137 + W 6 + out 410 MB)
    (1, 4096, 1024, 1, 0): 149.58e6 + 496.06e6,     # DINOv2 fc1   (A 137 + W 8 + out 547 MB)
    (2, 1024, 4096, 0, 1): 928.53e6 + 378.77e6,     # DINOv2 fc2   (A 547 + W 8 + resid 274 r + 274 w + bf16 copy 137 MB)
    (2, 1024, 1024, 0, 1): 415.44e6 + 362.06e6,     # DINOv2 proj  (A 137 + W 2 + resid 274 r + 274 w + bf16 copy 137 MB)
    (1, 4352, 1152, 1, 0): 167.40e6 + 524.24e6,     # SigLIP fc1   (A 151 + W 10 + out 570 MB)
    (2, 1152, 4352, 0, 1): 948.67e6 + 425.22e6,     # SigLIP fc2   (A 570 + W 10 + resid 302 r + 302 w + bf16 copy 151 MB)
}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.rows, self.proc, self.thread, self.idx = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_images_per_s(steps: int, warmup: int, images_per_step: int = 1):
    """fp32 restatement of the reference PyTorch path (oracle/vit_oracle.py), all host threads, executing all
    24+27 blocks like timm 0.9.10 does.  Returns (images/s, ms/step, cores)."""
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import (make_projector_state_dict, make_vit_state_dict, normalize_frames,
                                         synthetic_frames)
    from oracle import vit_oracle   # allowed here: bench.py's cpu_baseline / --impl reference leg only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dsd = make_vit_state_dict(DINOV2_L14_REG4, seed=1234, init="timm")
    ssd = make_vit_state_dict(SIGLIP_SO400M_14, seed=1235, init="timm")
    psd = make_projector_state_dict(seed=4321, init="timm")
    px = normalize_frames(synthetic_frames(images_per_step, seed=0))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vit_oracle.featurize_project(dsd, DINOV2_L14_REG4, ssd, SIGLIP_SO400M_14, psd, px, run_all_blocks=True)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return images_per_step * len(times) / total, 1e3 * total / len(times), cores


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    ips, ms, cores = cpu_reference_images_per_s(steps, warmup, images_per_step=1)
    sample = f"{steps} timed + {warmup} warm-up forwards of 1 image (fp32, all 24+27 blocks + projector)"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path; batch 1 per step"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def build_encoder(device):
    import bridgelang_b200 as blb
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict

    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    bb.dino_featurizer.load_state_dict(make_vit_state_dict(DINOV2_L14_REG4, seed=1234, init="timm"))
    bb.siglip_featurizer.load_state_dict(make_vit_state_dict(SIGLIP_SO400M_14, seed=1235, init="timm"))
    proj = blb.FusedMLPProjector(bb.embed_dim, 4096)
    proj.load_state_dict(make_projector_state_dict(seed=4321, init="timm"))
    return blb.VisualPrefixEncoder(bb, proj).to(device)


def run_gpu(args) -> None:
    import torch.distributed as dist

    from bridgelang_b200 import ops
    from bridgelang_b200.build import build_library
    from bridgelang_b200.config import fused_flops_per_image
    from bridgelang_b200.pipeline import gather_prefixes
    from bridgelang_b200.weights import normalize_frames, synthetic_frames

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        build_library()
    if distributed:
        dist.barrier()

    B = args.batch
    global_batch = B * world
    enc = build_encoder(device)
    # this rank's contiguous slice of the global synthetic batch (different frames per rank)
    frames = synthetic_frames(B, seed=1000 + rank)
    px_host = {k: v.to(torch.bfloat16).pin_memory() for k, v in normalize_frames(frames).items()}
    frames_host = frames.contiguous().pin_memory()
    px_dev = {k: v.to(device, non_blocking=True) for k, v in px_host.items()}
    torch.cuda.synchronize()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        out = enc(px_dev)
        if args.gather and distributed:
            out = gather_prefixes(out, global_batch)
        return out

    def endless(x):
        while True:
            yield x

    # public serving API: VisualPrefixEncoder.stream() double-buffers the H2D copy of batch i+1 under the encode of
    # batch i; every step still copies its own full batch from pinned host memory and reads its own result back
    e2e_stream = enc.stream(endless(px_host))
    e2e_stream_u8 = enc.stream(endless(frames_host), uint8=True)

    def step_e2e_uint8():
        # SURVEY §8f.2 variant: the host hands over the resized uint8 frame (4x fewer H2D bytes); ToTensor + both
        # Normalizes run in one device kernel.  Reported next to `e2e`, never instead of it.
        out = next(e2e_stream_u8)
        if args.gather and distributed:
            out = gather_prefixes(out, global_batch)
        return out.mean(dim=(1, 2), dtype=torch.float32).cpu()

    def step_e2e():
        out = next(e2e_stream)
        if args.gather and distributed:
            out = gather_prefixes(out, global_batch)
        digest = out.mean(dim=(1, 2), dtype=torch.float32)   # one fp32 per image (fp32 accumulation, no 1 GB copy)
        return digest.cpu()                              # D2H read of the step's result

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- headline: device-resident inputs -----------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    total_ms = timed(step_resident, args.steps, args.warmup)
    launches_per_step = (ops.launch_count() - n0) // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = global_batch / (ms_per_step * 1e-3)

    # ---- e2e: host frames → H2D → forward → D2H digest -------------------------------------------------------
    e2e_steps = max(2, args.steps)
    e2e_ms = timed(step_e2e, e2e_steps, 2) / e2e_steps
    e2e_value = global_batch / (e2e_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in px_host.values())
    d2h = B * 4
    e2e_u8_ms = timed(step_e2e_uint8, max(2, min(args.steps, 5)), 2) / max(2, min(args.steps, 5))

    # ---- roofline of the dominant kernel (instrumented pass, not the headline number) ------------------------
    roof = None
    if rank == 0:
        peaks, src = load_peaks()
        ops.timing_enable(True)
        ops.timing_reset()
        torch.cuda.synchronize()
        for _ in range(2):
            enc(px_dev)              # rank-local on purpose: no collective inside a rank-0-only block
        torch.cuda.synchronize()
        cats = ops.timing_collect()
        recs = ops.timing_records(8192)
        ops.timing_enable(False)
        ops.timing_reset()
        g = cats["gemm"]
        family = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        step_tflops = fused_flops_per_image() * B / (ms_per_step * 1e-3) / 1e12
        # dominant kernel = the GEMM shape with the largest share of the step; its per-launch numbers
        shapes = {}
        for r in recs:
            if r["cat"] != "gemm":
                continue
            key = (r["mode"], r["N"], r["K"], r["ln_folded"], r["emits_stats"])
            a = shapes.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += r["ms"]
            a[2] += r["work"]
        top = max(shapes.items(), key=lambda kv: kv[1][1])
        (mode, Nn, Kk, lnf, st), (cnt, tms, twork) = top
        mode_name = {0: "bias", 1: "bias+GELU", 2: "LayerScale+residual", 3: "patch"}[mode]
        achieved = twork / (tms * 1e-3) / 1e12
        roof = {
            "bound": "tensor",
            "kernel": f"gemm_bf16_kernel tcgen05 cta_group::2, epilogue {mode_name}"
                      f"{' + folded LayerNorm' if lnf else ''}{' + stats/bf16-copy' if st else ''}, "
                      f"M={int(round(twork / cnt / (2.0 * Nn * Kk)))} N={Nn} K={Kk}",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
            "flops_per_launch": twork / cnt, "avg_launch_us": 1e3 * tms / cnt, "launches_per_step": cnt // 2,
            "share_of_step": (tms / 2) / ms_per_step,
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((mode, Nn, Kk, lnf, st)),
            "traffic_source": "profiles/r01_ncu_full_c_kernels.md (ncu --set full, dram__bytes_read+write per launch)",
            "gemm_family": {"achieved": family, "frac": family / peak, "note": "all tcgen05 GEMM launches of the step"},
            "gemm_ms_per_step": g["ms"] / 2, "gemm_launches_per_step": g["launches"] // 2,
            "whole_step": {"achieved": step_tflops, "frac_of_sustained": step_tflops / peak,
                           "frac_of_burst": step_tflops / float(peaks["bf16_tflops"]),
                           "flops_per_image": fused_flops_per_image()},
            "breakdown_ms_per_step": {k: v["ms"] / 2 for k, v in cats.items()},
            "attention_tflops": (cats["attention"]["work"] / (cats["attention"]["ms"] * 1e-3) / 1e12
                                 if cats["attention"]["ms"] > 0 else None),
            "layernorm_gbs": (cats["layernorm"]["work"] / (cats["layernorm"]["ms"] * 1e-3) / 1e9
                              if cats["layernorm"]["ms"] > 0 else None),
        }

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, ms, cores = cpu_reference_images_per_s(steps=20, warmup=2, images_per_step=1)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "20 timed + 2 warm-up forwards of 1 image, fp32 oracle restatement incl. the wasted last "
                         "block of each tower (420.15 GFLOP/image)", "ms_per_image": ms}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": global_batch,
                       "parallelism": f"dp{world}", "weights": "random timm-init, seed 1234",
                       "cache": "no L2 flush needed: per-step working set (1.6 GB weights + >2 GB activations) "
                                "exceeds the 126 MB L2",
                       "collective": "nccl all-gather of prefixes" if (args.gather and distributed) else "none"},
            "clocks": clocks,
            "e2e_uint8": {"value": global_batch / (e2e_u8_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_u8_ms,
                          "h2d_bytes_per_step": frames_host.numel(), "d2h_bytes_per_step": d2h,
                          "note": "extra (SURVEY 8f.2): resized uint8 HWC frames from pinned host memory; ToTensor + "
                                  "both Normalizes in one device kernel (bit-identical to the host transform)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--gather", action="store_true", help="include the NCCL all-gather of projected prefixes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
