import torch, time
n=2_500_000_000
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for chunk in (n, 64<<20, 256<<20):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3):
        for lo in range(0,n,chunk):
            d[lo:lo+chunk].copy_(h[lo:lo+chunk],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print('1D chunk',chunk>>20,'MB', n/dt/1e9,'GB/s')
