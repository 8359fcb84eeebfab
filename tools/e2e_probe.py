"""e2e_probe.py — H2D staging probe (development tool): serial vs overlapped copies, VisualPrefixEncoder.stream(),
and the cost of the per-step digest.  Finding of round 1: a host tensor whose strides differ from the device staging
buffer turns `copy_(non_blocking=True)` into a host re-layout + synchronous copy (+15 ms per 154 MB batch)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import build_encoder
from bridgelang_b200.weights import normalize_frames, synthetic_frames

dev = torch.device("cuda", 0)
enc = build_encoder(dev)
frames = synthetic_frames(256, seed=1000)
px_host = {k: v.to(torch.bfloat16).pin_memory() for k, v in normalize_frames(frames).items()}
px_dev = {k: v.to(dev) for k, v in px_host.items()}
slot = {k: torch.empty_like(v) for k, v in px_dev.items()}
copier = torch.cuda.Stream(dev)


def ev_time(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def h2d_same_stream():
    for k, v in px_host.items():
        slot[k].copy_(v, non_blocking=True)


def h2d_copier():
    with torch.cuda.stream(copier):
        for k, v in px_host.items():
            slot[k].copy_(v, non_blocking=True)
    torch.cuda.current_stream().wait_stream(copier)


def fwd():
    enc(px_dev)


def both_overlapped():
    with torch.cuda.stream(copier):
        for k, v in px_host.items():
            slot[k].copy_(v, non_blocking=True)
    enc(px_dev)
    torch.cuda.current_stream().wait_stream(copier)


def serial():
    h2d_same_stream()
    enc(slot)


print("pinned:", {k: v.is_pinned() for k, v in px_host.items()})
print(f"h2d same stream   {ev_time(h2d_same_stream):8.2f} ms")
print(f"h2d copier stream {ev_time(h2d_copier):8.2f} ms")
print(f"forward           {ev_time(fwd):8.2f} ms")
print(f"serial h2d+fwd    {ev_time(serial):8.2f} ms")
print(f"overlapped        {ev_time(both_overlapped):8.2f} ms")
t0 = time.perf_counter(); enc(px_dev); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host launch time of one forward {1e3*(t1-t0):.2f} ms, device drain {1e3*(t2-t1):.2f} ms")


def endless(x):
    while True:
        yield x


g = enc.stream(endless(px_host))
print(f"stream() only           {ev_time(lambda: next(g)):8.2f} ms")
print(f"stream() only, again    {ev_time(lambda: next(g)):8.2f} ms")
print(f"stream() only, n=15     {ev_time(lambda: next(g), n=15):8.2f} ms")
import os
os.environ["X"] = "1"
def held():
    held.o = enc(px_dev)          # keep the previous output alive like the generator does
print(f"forward, output held    {ev_time(held):8.2f} ms")
held.q = []
def held2():
    held.q.append(enc(px_dev)); held.q = held.q[-2:]
print(f"forward, 2 outputs held {ev_time(held2):8.2f} ms")
print(f"stream() + digest.cpu() {ev_time(lambda: next(g).float().mean(dim=(1, 2)).cpu()):8.2f} ms")
print(f"forward + digest.cpu()  {ev_time(lambda: enc(px_dev).float().mean(dim=(1, 2)).cpu()):8.2f} ms")
print(f"serial + digest.cpu()   {ev_time(lambda: enc({k: v.to(dev, non_blocking=True) for k, v in px_host.items()}).float().mean(dim=(1, 2)).cpu()):8.2f} ms")
out = enc(px_dev)
print(f"digest only             {ev_time(lambda: out.float().mean(dim=(1, 2)).cpu()):8.2f} ms")
print(f"digest (no fp32 copy)   {ev_time(lambda: out.mean(dim=(1, 2), dtype=torch.float32).cpu()):8.2f} ms")
print(torch.cuda.memory_summary(abbreviated=True)[:1200])

