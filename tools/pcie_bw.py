"""Pinned host <-> device copy bandwidth on this box (context for the e2e numbers)."""
import time, torch
n = 155_478_862
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print("%s %.1f MB in %.2f ms = %.1f GB/s" % (name, n / 1e6, dt * 1e3, n / dt / 1e9))
