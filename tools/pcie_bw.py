#!/usr/bin/env python3
"""Pinned-memory PCIe bandwidth of the box: H2D alone, D2H alone, both directions at once
(the ceiling of bench.py's e2e number)."""
import torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return reps * n * (int(h2d) + int(d2h)) / ms / 1e6


run(True, True, 2)
print("H2D alone  %.1f GB/s" % run(True, False))
print("D2H alone  %.1f GB/s" % run(False, True))
print("both       %.1f GB/s (sum of the two directions)" % run(True, True))
