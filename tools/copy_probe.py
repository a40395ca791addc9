import time, torch
dev = torch.device("cuda", 0)
n = 20200 * 4096
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for name, fn in [("h2d pinned", lambda: d.copy_(h, non_blocking=True)), ("d2h into existing pinned", lambda: h.copy_(d, non_blocking=True)),
                 ("h.to(dev)", lambda: h.to(dev, non_blocking=True)),
                 ("alloc pinned + d2h", lambda: torch.empty(n, dtype=torch.float64, pin_memory=True).copy_(d, non_blocking=True)),
                 ("d.cpu() pageable", lambda: d.cpu())]:
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:28s} rep{rep}: {dt*1e3:8.1f} ms  {n*8/dt/1e9:6.1f} GB/s")
        del r
