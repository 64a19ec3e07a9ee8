import torch, time
d = torch.device("cuda:0")
for mb in (2, 8, 25, 100):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    g = torch.empty(n, dtype=torch.uint8, device=d)
    for name, fn in (("D2H", lambda: h.copy_(g, non_blocking=True)), ("H2D", lambda: g.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"{name} {mb:4d} MB pinned: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
