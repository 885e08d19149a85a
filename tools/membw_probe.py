#!/usr/bin/env python
"""Plain read / write / copy bandwidth of the box (torch kernels, CUDA events): what a pure stream achieves,
to put the head/tail kernels' achieved DRAM rates (profiles/r02_bigfft_ncu.txt) next to."""
import torch
n = 512 * 1024 * 1024
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.view(torch.float32).fill_(1.0)); print(f"write 512 MiB (fill_ f32): {ms*1e3:7.1f} us  {n/ms/1e6:7.0f} GB/s")
ms = t(lambda: a.view(torch.int64).sum());       print(f"read  512 MiB (sum i64) : {ms*1e3:7.1f} us  {n/ms/1e6:7.0f} GB/s")
ms = t(lambda: b.copy_(a));                      print(f"copy  512 MiB           : {ms*1e3:7.1f} us  {2*n/ms/1e6:7.0f} GB/s (read+write)")
