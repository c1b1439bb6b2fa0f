"""H2D / D2H bandwidth of pinned host memory on this box (torch is only used here, as a probe)."""
import time
import torch
n = 300 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, f in (('H2D', lambda: d.copy_(h, non_blocking=True)), ('D2H', lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print('%s 300 MB: %.2f ms  %.1f GB/s' % (name, dt * 1e3, n / dt / 1e9))
