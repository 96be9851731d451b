#!/usr/bin/env python
"""Developer helper (GPU box): pinned-memory H2D / D2H bandwidth, alone and both directions at once."""
import time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d1.copy_(h1, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print({"h2d_GBs": n / a / 1e9, "d2h_GBs": n / b / 1e9, "both_each_GBs": n / c / 1e9})
