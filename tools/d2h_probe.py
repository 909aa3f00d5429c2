#!/usr/bin/env python
"""Pinned device-to-host copy rate of this box (the ceiling of bench.py's e2e number: 5.9 GB of Surface Temperature per step)."""
import time

import torch

n = 5888802816 // 8
d = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
h = torch.empty(n, dtype=torch.float64).pin_memory()
for _ in range(2):
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print("pinned D2H of %.1f GB: %.1f ms, %.1f GB/s" % (n * 8 / 1e9, dt * 1e3, n * 8 / dt / 1e9))
