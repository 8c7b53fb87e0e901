#!/usr/bin/env python
"""N-GPU check of the NVLink peer-memory all-reduce against NCCL:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_allreduce_check.py
Every rank sums random partial phenotypes 200 times (fresh data each time: a stale read or a race
between epochs shows up as a mismatch), then both are timed with CUDA events."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simu_b200 import capi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
bad = 0
for n, k in ((100_000, 2), (100_000, 1), (500_000, 2), (1, 1), (12_345, 2), (33_333, 1), (7, 2)):
    count = n * k
    ctx = capi.Context(max(n, 4), device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.peer_init(rank, world, count))
    ctx.peer_connect(handles)
    g = torch.Generator(device=dev).manual_seed(1000 * rank + n)
    worst = 0.0
    for it in range(200):
        y = torch.randn(count, dtype=torch.float64, device=dev, generator=g)
        ref = y.clone()
        dist.all_reduce(ref)
        ctx.peer_allreduce_dev(y.data_ptr(), count)
        worst = max(worst, float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-300)))
        # identical bits on every rank (rank-ordered sum)
        chk = y.view(torch.int64).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if lo.item() != hi.item():
            bad += 1
    y = torch.randn(count, dtype=torch.float64, device=dev, generator=g)

    def timed(fn, reps=200):
        for _ in range(20):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # the exchange replayed from a CUDA graph (its epoch lives in device memory, so the launch parameters never change):
    # fresh data before every replay, compared with NCCL as above
    yg = torch.zeros(count, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        ctx.peer_allreduce_dev(yg.data_ptr(), count)
    for it in range(30):
        fresh = torch.randn(count, dtype=torch.float64, device=dev, generator=g)
        ref = fresh.clone()
        dist.all_reduce(ref)
        yg.copy_(fresh)
        graph.replay()
        worst = max(worst, float((yg - ref).abs().max() / ref.abs().max().clamp_min(1e-300)))
    torch.cuda.synchronize()
    del graph
    t_peer = timed(lambda: ctx.peer_allreduce_dev(y.data_ptr(), count))
    t_nccl = timed(lambda: dist.all_reduce(y))
    if worst > 1e-12:
        bad += 1
    if rank == 0:
        print(f"peer_allreduce world={world} count={count}: max rel diff vs NCCL {worst:.1e}, rank-identical bits "
              f"{'yes' if not bad else 'NO'}; peer {t_peer:.1f} us, NCCL {t_nccl:.1f} us per call", flush=True)
    torch.cuda.synchronize()
    ctx.close()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
