import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simu_b200 import capi
for n in (100_000, 500_000):
    ctx = capi.Context(n)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
    y = torch.randn(n, dtype=torch.float64, device='cuda'); idx = torch.empty(n, dtype=torch.int32, device='cuda')
    for _ in range(3): ctx.liability_order_dev(y.data_ptr(), n, idx.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record(s)
    for _ in range(20): ctx.liability_order_dev(y.data_ptr(), n, idx.data_ptr())
    e1.record(s); torch.cuda.synchronize()
    ok = bool(torch.equal(y[idx.long()], torch.sort(y).values))
    print(n, 'sort ms', e0.elapsed_time(e1) / 20, 'launches/sort', (ctx.launches - l0) / 20, 'ok', ok)
    ctx.close()
