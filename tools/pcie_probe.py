#!/usr/bin/env python
"""Box probe: pinned host <-> device copy bandwidth, one direction and both at once (the bound of bench.py's e2e)."""
import time
import torch
n = 56623104
h1 = torch.empty(n, dtype=torch.float64).pin_memory()
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda")
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


gb = n * 8 / 1e9
a = t(lambda: d1.copy_(h1, non_blocking=True))
b = t(lambda: h2.copy_(d2, non_blocking=True))


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


c = t(both)
print(f"H2D {gb / a:.1f} GB/s ({a * 1e3:.2f} ms)  D2H {gb / b:.1f} GB/s ({b * 1e3:.2f} ms)  both at once {c * 1e3:.2f} ms "
      f"({2 * gb / c:.1f} GB/s total)")
