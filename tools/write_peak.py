#!/usr/bin/env python
"""Pure-write HBM bandwidth on this GPU (what a write-only stream can reach), to put the
observation-write roofline in context: torch fill_ / zero_ / copy_ over 1.6 GB, best of 20."""
import torch

n = 1610612736
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
xi = x.view(torch.int32)


def best(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return min(t), sorted(t)[len(t) // 2]


for name, fn, nbytes in [("fill_u8", lambda: x.fill_(7), n), ("fill_i32", lambda: xi.fill_(7), n),
                         ("zero_ (memset)", lambda: x.zero_(), n), ("copy_ (r+w)", lambda: y.copy_(x), 2 * n)]:
    lo, med = best(fn)
    print(f"{name:16s} best {nbytes / lo / 1e6:8.1f} GB/s   median {nbytes / med / 1e6:8.1f} GB/s")
