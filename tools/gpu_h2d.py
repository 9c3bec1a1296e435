#!/usr/bin/env python
"""Pinned host -> device copy rate of the step's 264 MB of inputs, alone and while the GPU is busy."""
import torch, time
x = torch.empty(66_000_000, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("H2D 264 MB alone: %.3f ms = %.1f GB/s" % (ms, 0.264 / ms * 1e3))
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
torch.cuda.synchronize()
with torch.cuda.stream(s):
    e0.record(s)
    for _ in range(10):
        d.copy_(x, non_blocking=True)
    e1.record(s)
for _ in range(40):
    a @ b
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("H2D 264 MB under a busy GPU: %.3f ms = %.1f GB/s" % (ms, 0.264 / ms * 1e3))
