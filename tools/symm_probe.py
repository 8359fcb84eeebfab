"""symm_probe.py — does torch.distributed._symmetric_memory work on this box?  Each rank pushes its shard into every
peer's buffer with plain device-to-device copies (copy engines, no SMs), then a barrier."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import torch.distributed._symmetric_memory as symm
n = 256 * 256 * 4096            # one shard: 512 MiB of bf16
buf = symm.empty((world, n), dtype=torch.bfloat16, device=f"cuda:{rank}")
hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
local = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device="cuda")
peers = [hdl.get_buffer(r, (world, n), torch.bfloat16) for r in range(world)]
copy_stream = torch.cuda.Stream()
def gather():
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev)
        for k in range(world):
            r = (rank + k) % world
            peers[r][rank].copy_(local, non_blocking=True)
        hdl.barrier(channel=0)
for _ in range(2):
    gather()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    gather()
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / 5
ok = all(bool((buf[r] == r + 1).all()) for r in range(world))
print(f"rank {rank}: ok={ok} {dt*1e3:.2f} ms per gather, {n*2*(world-1)/dt/1e9:.0f} GB/s out per rank", flush=True)
dist.destroy_process_group()
