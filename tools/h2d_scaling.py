"""Probe (torchrun): host->device bandwidth per rank when all ranks copy at once, with the
process left where the OS put it vs bound to the CPU set NVML reports for its GPU (NUMA-local
pinned memory)."""
import os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
if mode == "affinity":
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        print(f"rank {rank}: gpu {local} affinity {len(cpus)} cpus: {cpus[:4]}..{cpus[-2:] if cpus else ''}", flush=True)
    except Exception as e:
        print(f"rank {rank}: affinity failed: {e}", flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
src = torch.empty(200 * 1024 * 1024, dtype=torch.uint8).pin_memory()
src.fill_(1)
dst = torch.empty_like(src, device=dev)
back = torch.empty(100 * 1024 * 1024, dtype=torch.uint8).pin_memory()
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = 20 * src.numel() / dt / 1e9
s2 = torch.cuda.Stream()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    dst.copy_(src, non_blocking=True)
    with torch.cuda.stream(s2):
        back.copy_(dst[: back.numel()], non_blocking=True)
torch.cuda.synchronize()
dt2 = time.perf_counter() - t0
print(f"[{mode}] rank {rank}/{world}: H2D {gbs:.1f} GB/s alone; with concurrent D2H: H2D {20*src.numel()/dt2/1e9:.1f} GB/s + D2H {20*back.numel()/dt2/1e9:.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
