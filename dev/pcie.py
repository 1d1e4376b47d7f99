import torch, time
dev=torch.device("cuda",0)
for mb in (1, 8, 25, 64, 256):
    n=mb*1024*1024//8
    h=torch.empty(n,dtype=torch.float64).pin_memory(); d=torch.empty(n,dtype=torch.float64,device=dev)
    s=torch.cuda.Stream(); s2=torch.cuda.Stream()
    for _ in range(3): d.copy_(h,non_blocking=True); h.copy_(d,non_blocking=True)
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); 
    for _ in range(5): d.copy_(h,non_blocking=True)
    b.record(); torch.cuda.synchronize(); t1=a.elapsed_time(b)/5
    a.record(); 
    for _ in range(5): h.copy_(d,non_blocking=True)
    b.record(); torch.cuda.synchronize(); t2=a.elapsed_time(b)/5
    # duplex
    h2=torch.empty(n,dtype=torch.float64).pin_memory(); d2=torch.empty(n,dtype=torch.float64,device=dev)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s): d.copy_(h,non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    torch.cuda.synchronize(); t3=(time.perf_counter()-t0)*1e3/5
    print("%4d MB  H2D %.2f ms %.1f GB/s | D2H %.2f ms %.1f GB/s | duplex %.2f ms"%(mb,t1,mb/1024/t1*1e3,t2,mb/1024/t2*1e3,t3))
