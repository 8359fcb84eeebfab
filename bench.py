#!/usr/bin/env python
"""bench.py — images/s of the prism-dinosiglip-224px featurize+project path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B | --global-batch G] [--tower dino|siglip]
                    [--gather] [--impl reference]

A step = one featurize+project forward over one batch of B=256 synthetic 224x224 frames per GPU (BASELINE config
"prism-dinosiglip-224px featurizer + projector, bf16, batch 256 synthetic 224px frames, 1xB200"); N>1 shards a
global batch of 256*N images by image (weak scaling, no collective in the path; --gather adds the optional NCCL
all-gather of the projected prefixes).  One JSON line is printed by rank 0:

  value      whole-job images/s with the normalized bf16 frames already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API (VisualPrefixEncoder.stream) from PINNED HOST frames: every step copies
             its own batch host→device AND its own full projected prefix [B,256,4096] device→host (pinned buffers,
             both copies overlapped with the next batch's encode on side streams) inside the timed region;
             e2e_digest is the variant whose consumer stays on the device (only a per-image fp32 digest comes back)
  --global-batch G   BASELINE configs[2]: G images sharded by image over the N ranks (2048 → 1024 / 512 / 256 per GPU)
  --tower T          BASELINE configs[3]: one tower alone (SigLIPViTBackbone / DinoV2ViTBackbone), B=256
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
    (2, 1024, 1024, 0, 1): 414.92e6 + 360.50e6,     # DINOv2 proj  (A 137 + W 2 + resid 274 r + 274 w + bf16 copy 137 MB)   [r02]
    (1, 4352, 1152, 1, 0): 167.39e6 + 521.80e6,     # SigLIP fc1   (A 151 + W 10 + out 570 MB)   [r02]
    (2, 1152, 4352, 0, 1): 939.50e6 + 424.58e6,     # SigLIP fc2   (A 570 + W 10 + resid 302 r + 302 w + bf16 copy 151 MB)   [r02]
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
    steps, warmup = max(1, args.steps), max(0, args.warmup)     # exactly what the driver asked for
    ips, ms, cores = cpu_reference_images_per_s(steps, warmup, images_per_step=1)
    sample = f"{steps} timed + {warmup} warm-up forwards of 1 image (fp32, all 24+27 blocks + projector)"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (fp32 oracle port, all 24+27 "
                   "blocks as timm executes them); each step = a bounded sample of the workload: 1 image; one host "
                   "process on all cores whatever --gpus says, so only the N=1 ratio is like for like"},
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


def build_single_tower(device, which: str):
    """BASELINE configs[3]: the reference's single-encoder backbones (siglip_vit.py:8-24, dinov2_vit.py:9-19)."""
    import bridgelang_b200 as blb
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_vit_state_dict

    if which == "dino":
        bb, cfg, seed = blb.DinoV2ViTBackbone("dinov2-vit-l", "resize-naive"), DINOV2_L14_REG4, 1234
    else:
        bb, cfg, seed = blb.SigLIPViTBackbone("siglip-vit-so400m", "resize-naive"), SIGLIP_SO400M_14, 1235
    bb.featurizer.load_state_dict(make_vit_state_dict(cfg, seed=seed, init="timm"))
    return bb.to(device), cfg


def run_gpu(args) -> None:
    import torch.distributed as dist

    from bridgelang_b200 import ops
    from bridgelang_b200.build import build_library
    from bridgelang_b200.config import fused_flops_per_image
    from bridgelang_b200.pipeline import PrefixGatherer, gather_prefixes, shard_bounds
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
        if args.gather and not args.gather_inline and args.gather_mode == "nccl" and args.nccl_max_ctas > 0:
            # the overlapped all-gather shares the SMs with the towers: a few CTAs move 3.5 GiB in the ~100 ms of a step
            # with room to spare, more only take SMs and HBM bandwidth from the GEMMs
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        build_library()
    if distributed:
        dist.barrier()

    # ---- workload: images per rank -----------------------------------------------------------------------------
    if args.global_batch:          # BASELINE configs[2]: a fixed global batch sharded by image (strong scaling)
        global_batch = args.global_batch
        lo, hi = shard_bounds(global_batch, rank, world)
        B, scaling = hi - lo, "strong"
    else:                          # default / driver contract: 256 images per GPU (weak scaling)
        B, global_batch, scaling = args.batch, args.batch * world, "weak"
    single = args.tower is not None
    if single:
        model, cfg = build_single_tower(device, args.tower)
        flops_per_image = cfg.flops_per_image()
        tower_key = args.tower
        metric = f"images/sec {'DINOv2 ViT-L/14-reg4' if args.tower == 'dino' else 'SigLIP SO400M/14'}-224px featurize bf16 b=256"
        workload = (f"single-encoder {'dinov2-vit-l' if args.tower == 'dino' else 'siglip-vit-so400m'} backbone, bf16, "
                    f"batch {B} synthetic 224px frames per GPU (BASELINE configs[3])")
    else:
        model = build_encoder(device)
        flops_per_image = fused_flops_per_image()
        metric, workload = METRIC, WORKLOAD
    # this rank's contiguous slice of the global synthetic batch (different frames per rank)
    frames = synthetic_frames(B, seed=1000 + rank)
    px_all = normalize_frames(frames)
    if single:
        px_host = px_all[tower_key].to(torch.bfloat16).pin_memory()
        px_dev = px_host.to(device, non_blocking=True)
    else:
        px_host = {k: v.to(torch.bfloat16).pin_memory() for k, v in px_all.items()}
        px_dev = {k: v.to(device, non_blocking=True) for k, v in px_host.items()}
    frames_host = frames.contiguous().pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # --gather: the NCCL all-gather of step i's prefixes runs on a communication stream under the towers of step i+1
    # (PrefixGatherer); --gather-inline keeps it on the compute stream (the round-1 behaviour, for the A/B)
    gatherer = (PrefixGatherer(global_batch, mode=args.gather_mode)
                if (args.gather and distributed and not single) else None)
    in_flight = [None]

    def step_resident():
        out = model(px_dev)
        if gatherer is not None:
            if args.gather_inline:
                out = gather_prefixes(out, global_batch)
            else:
                handle = gatherer.launch(out)
                if in_flight[0] is not None:
                    gatherer.wait(in_flight[0])      # the consumer takes result i-1 while the gather of i is in flight
                in_flight[0] = handle
        return out

    def endless(x):
        while True:
            yield x

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

    # ---- e2e: the public serving call from pinned host frames; H2D of the inputs and D2H of the FULL result inside ----
    e2e = e2e_digest = e2e_u8 = None
    if not single:
        e2e_steps = max(2, args.steps)
        pending = []

        # VisualPrefixEncoder.stream(to_host=True): batch i+1's H2D copy and batch i-1's D2H copy run on side streams
        # under the encode of batch i; every step still moves its own 154 MB in and its own [B,256,4096] prefix out
        e2e_stream_full = model.stream(endless(px_host), to_host=True)

        def step_e2e_full():
            host_out, done = next(e2e_stream_full)
            pending.append(done)
            if len(pending) > 1:
                pending.pop(0).synchronize()       # the consumer reads result i-1 while step i runs

        def drain():
            while pending:
                pending.pop(0).synchronize()

        # warm-up: both pinned host buffers, both staging slots AND the third device result block (two results are
        # still being drained when the next one is allocated) exist before the clock starts
        for _ in range(max(3, args.warmup)):
            step_e2e_full()
        drain()
        barrier()
        t0 = time.perf_counter()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for _ in range(e2e_steps):
            step_e2e_full()
        drain()                                    # the last result is on the host before the clock stops
        e_ev.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = s_ev.elapsed_time(e_ev)
        tms = torch.tensor([max(dev_ms, wall_ms)], device=device, dtype=torch.float64)
        if distributed:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = tms.item() / e2e_steps
        h2d = sum(v.numel() * v.element_size() for v in px_host.values())
        d2h_full = B * 256 * 4096 * 2
        e2e = {"value": global_batch / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h_full, "ms_per_step": e2e_ms,
               "device_ms_per_step": dev_ms / e2e_steps, "host_ms_per_step": wall_ms / e2e_steps,
               "note": "VisualPrefixEncoder.stream(to_host=True): pinned host frames in, the full bf16 prefix "
                       "[B,256,4096] back in pinned host memory, every step; timed by max(CUDA events, host clock)"}

        e2e_stream = model.stream(endless(px_host))
        e2e_stream_u8 = model.stream(endless(frames_host), uint8=True)

        def step_e2e_digest():
            out = next(e2e_stream)
            if args.gather and distributed:
                out = gather_prefixes(out, global_batch)
            return out.mean(dim=(1, 2), dtype=torch.float32).cpu()   # one fp32 per image: the consumer is on-device

        def step_e2e_uint8():
            # SURVEY §8f.2 variant: the host hands over the resized uint8 frame (4x fewer H2D bytes); ToTensor + both
            # Normalizes run in one device kernel.  Reported next to `e2e`, never instead of it.
            out = next(e2e_stream_u8)
            if args.gather and distributed:
                out = gather_prefixes(out, global_batch)
            return out.mean(dim=(1, 2), dtype=torch.float32).cpu()

        dig_ms = timed(step_e2e_digest, e2e_steps, 3) / e2e_steps
        e2e_digest = {"value": global_batch / (dig_ms * 1e-3), "unit": UNIT, "ms_per_step": dig_ms,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 4,
                      "note": "same, but the consumer (the LLM) stays on the device: only a per-image fp32 digest is read back"}
        n8 = max(2, min(args.steps, 5))
        u8_ms = timed(step_e2e_uint8, n8, 3) / n8
        e2e_u8 = {"value": global_batch / (u8_ms * 1e-3), "unit": UNIT, "ms_per_step": u8_ms,
                  "h2d_bytes_per_step": frames_host.numel(), "d2h_bytes_per_step": B * 4,
                  "note": "extra (SURVEY 8f.2): resized uint8 HWC frames from pinned host memory; ToTensor + both "
                          "Normalizes in one device kernel (bit-identical to the host transform)"}
    else:
        # single tower: H2D of the tower's frames, D2H of its [B,256,D] features, plain forward per step
        out_host = torch.empty((B, 256, cfg.dim), dtype=torch.bfloat16, pin_memory=True)
        dev_in = torch.empty_like(px_host, device=device)

        def step_e2e_single():
            dev_in.copy_(px_host, non_blocking=True)
            out_host.copy_(model(dev_in), non_blocking=True)

        e2e_steps = max(2, args.steps)
        e2e_ms = timed(step_e2e_single, e2e_steps, 3) / e2e_steps
        e2e = {"value": global_batch / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": px_host.numel() * 2, "d2h_bytes_per_step": out_host.numel() * 2}

    # ---- roofline of the dominant kernel (instrumented pass, not the headline number) ------------------------
    roof = None
    if rank == 0:
        peaks, src = load_peaks()
        ops.timing_enable(True)
        ops.timing_reset()
        torch.cuda.synchronize()
        for _ in range(2):
            model(px_dev)            # rank-local on purpose: no collective inside a rank-0-only block
        torch.cuda.synchronize()
        cats = ops.timing_collect()
        recs = ops.timing_records(8192)
        ops.timing_enable(False)
        ops.timing_reset()
        g = cats["gemm"]
        family = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        step_tflops = flops_per_image * B / (ms_per_step * 1e-3) / 1e12
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
        (mode, Nn, Kk, lnf, st), (cnt, tms_, twork) = top
        mode_name = {0: "bias", 1: "bias+GELU", 2: "LayerScale+residual", 3: "patch"}[mode]
        achieved = twork / (tms_ * 1e-3) / 1e12
        att = cats["attention"]
        sm_mhz = (clocks or {}).get("sm_mhz") or 0.0
        roof = {
            "bound": "tensor",
            "kernel": f"gemm_bf16_kernel tcgen05 cta_group::2, epilogue {mode_name}"
                      f"{' + folded LayerNorm' if lnf else ''}{' + stats/bf16-copy' if st else ''}, "
                      f"N={Nn} K={Kk} (tile-padded shape; flops_per_launch is the unpadded algorithmic 2·M·N·K)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
            "flops_per_launch": twork / cnt, "avg_launch_us": 1e3 * tms_ / cnt, "launches_per_step": cnt // 2,
            "share_of_step": (tms_ / 2) / ms_per_step,
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((mode, Nn, Kk, lnf, st)) if B == 256 else None,
            "traffic_source": "profiles/r02_ncu_gemm.md / r01_ncu_full_c_kernels.md (ncu --set full, dram__bytes_read+write per launch, B=256)",
            "gemm_family": {"achieved": family, "frac": family / peak, "note": "all tcgen05 GEMM launches of the step"},
            "gemm_ms_per_step": g["ms"] / 2, "gemm_launches_per_step": g["launches"] // 2,
            "whole_step": {"achieved": step_tflops, "frac_of_sustained": step_tflops / peak,
                           "frac_of_burst": step_tflops / float(peaks["bf16_tflops"]),
                           "flops_per_image": flops_per_image},
            "breakdown_ms_per_step": {k: v["ms"] / 2 for k, v in cats.items()},
            "attention": {
                "bound": "mufu", "kernel": "attention_tc_kernel (tcgen05 S/PV, softmax exp2 on MUFU.EX2; DINOv2 tail rows as a third tile of the same kernel)",
                "tflops": att["work"] / (att["ms"] * 1e-3) / 1e12 if att["ms"] > 0 else None,
                "ms_per_step": att["ms"] / 2, "launches_per_step": att["launches"] // 2,
                "floor": "16 exp2/clk/SM x 148 SMs (measured: profiles/r02_mufu_bench.log) = 2368 exp2/clk; one exp2 "
                         "per score → at the sampled SM clock the floor is scores / (2368 x clock)",
                "sm_mhz": sm_mhz,
            },
            "attention_tflops": (att["work"] / (att["ms"] * 1e-3) / 1e12 if att["ms"] > 0 else None),
            "layernorm_gbs": (cats["layernorm"]["work"] / (cats["layernorm"]["ms"] * 1e-3) / 1e9
                              if cats["layernorm"]["ms"] > 0 else None),
        }
        if att["ms"] > 0 and sm_mhz:
            # attention is bound by the MUFU pipe, not the tensor pipe: one exp2 per (query, key) score.  Scores per
            # image: DINOv2 23 blocks x 16 heads x 261², SigLIP 26 x 16 x 256² (whichever towers ran)
            from bridgelang_b200.config import DINOV2_L14_REG4 as _D, SIGLIP_SO400M_14 as _S
            per_img = 0
            for c in ((_D, _S) if not single else (cfg,)):
                per_img += c.n_needed_blocks * c.heads * c.tokens * c.tokens
            floor_ms = per_img * B / (2368.0 * sm_mhz * 1e6) * 1e3
            roof["attention"].update(exp2_per_step=per_img * B, mufu_floor_ms=floor_ms,
                                     frac=floor_ms / (att["ms"] / 2))
            # ... and it is about as close to the HBM bound: every launch reads the packed qkv tensor once and writes
            # the bf16 output once (algorithmic bytes 8·B·T·D per launch; ncu dram bytes agree within 3 %,
            # profiles/r02_ncu_attention.md) — round 2's diagnostic build without any exp2 is only 14 % faster
            hbm_bytes = 0
            for c in ((_D, _S) if not single else (cfg,)):
                hbm_bytes += c.n_needed_blocks * 8 * B * c.tokens * c.dim
            hbm_peak = float(peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6538.0)))
            hbm_floor_ms = hbm_bytes / (hbm_peak * 1e9) * 1e3
            roof["attention"].update(hbm_bytes_per_step=hbm_bytes, hbm_gbs=hbm_bytes / (att["ms"] / 2 * 1e-3) / 1e9,
                                     hbm_peak_gbs=hbm_peak, hbm_floor_ms=hbm_floor_ms,
                                     hbm_frac=hbm_floor_ms / (att["ms"] / 2))
            roof["attention"]["bound"] = "mufu + hbm (the two floors are within 20 % of each other)"

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not single:
        ips, ms, cores = cpu_reference_images_per_s(steps=20, warmup=2, images_per_step=1)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "20 timed + 2 warm-up forwards of 1 image, fp32 oracle restatement incl. the wasted last "
                         "block of each tower (420.15 GFLOP/image)", "ms_per_image": ms}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "global_batch": global_batch,
                       "parallelism": f"dp{world}", "weights": "random timm-init, seed 1234",
                       "cache": "no L2 flush needed: per-step working set (1.6 GB weights + >2 GB activations) "
                                "exceeds the 126 MB L2",
                       "collective": ("none" if gatherer is None else
                                      ("nccl all-gather of prefixes, in line" if args.gather_inline else
                                       ("peer-memory push (copy engines over NVLink)" if args.gather_mode == "p2p"
                                        else "nccl all-gather") + " of prefixes, overlapped with the next step's towers") +
                                      f", {2 * 256 * 4096 * (global_batch - B)} bytes received per rank per step")},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        if e2e_digest is not None:
            line["e2e_digest"] = e2e_digest
            line["e2e_uint8"] = e2e_u8
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (weak scaling, the default)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="BASELINE configs[2]: fixed global batch sharded by image over the ranks (e.g. 2048)")
    ap.add_argument("--tower", default=None, choices=["dino", "siglip"],
                    help="BASELINE configs[3]: one single-encoder backbone instead of the fused featurizer + projector")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--gather", action="store_true", help="include the NCCL all-gather of projected prefixes")
    ap.add_argument("--gather-inline", action="store_true", help="with --gather: keep the collective on the compute stream")
    ap.add_argument("--gather-mode", default="p2p", choices=["p2p", "nccl"],
                    help="overlapped --gather transport: peer-memory pushes on the copy engines (default) or NCCL")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="--gather-mode nccl: NCCL_MAX_CTAS for the communicator (0 = NCCL's default; measured: fewer is worse)")
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
