#!/usr/bin/env python
"""Host -> device copy ceiling of the box, per GPU and in aggregate, with every rank copying AT THE SAME TIME:
    python tools/h2d_probe.py                                                     (one GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py
Each rank copies 2.5 GB (config 2's causal rows) from pinned host memory in 64 MB pieces, once with the pinned
buffer allocated wherever the process happens to run and once after binding the process to the CPUs NVML
names as closest to its GPU (nvmlDeviceSetCpuAffinity) BEFORE the allocation, so that the pages are first
touched on the GPU's NUMA node.  The e2e leg of bench.py is bound by exactly this copy."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, chunk = 2_500_000_000, 64 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def measure(tag):
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h[::4096] = 1                                          # touch every page from this process
    for lo in range(0, n, chunk):
        d[lo:lo + chunk].copy_(h[lo:lo + chunk], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(3):
        for lo in range(0, n, chunk):
            d[lo:lo + chunk].copy_(h[lo:lo + chunk], non_blocking=True)
    torch.cuda.synchronize()
    gbs = 3 * n / (time.perf_counter() - t) / 1e9
    v = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    lo_, hi_, sm = v.clone(), v.clone(), v.clone()
    if world > 1:
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX); dist.all_reduce(sm)
    del h
    return {"placement": tag, "ranks": world, "per_gpu_gbs_min": round(lo_.item(), 1), "per_gpu_gbs_max": round(hi_.item(), 1),
            "aggregate_gbs": round(sm.item(), 1)}


out = [measure("default")]
try:
    import pynvml
    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    out.append(dict(measure("bound to the GPU's CPUs (NVML) before allocating"), cpus=len(os.sched_getaffinity(0))))
except Exception as e:                                       # noqa: BLE001
    out.append({"placement": "nvml affinity unavailable", "error": repr(e)})
if rank == 0:
    print(json.dumps({"h2d_probe": out, "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
