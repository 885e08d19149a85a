#!/usr/bin/env python
"""Is the fused kernel bound by memory at all?  The same kernel over 74 / 148 / 296 / 550 block pairs: 74 blocks
(78 MB of raw bytes) stay resident in the 126 MB L2 from one launch to the next, 550 (577 MB) stream from HBM.
If the time per block does not depend on where the bytes come from, the kernel is not memory-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

S, N = 262144, 4096
for NB in (74, 148, 296, 550):
    raw0, raw1 = synth.tiled_recording(NB, S, base_blocks=4)
    d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
    eng = FxEngine(S, N, 4, max_blocks=NB)
    out = (torch.empty((NB, N), dtype=torch.complex64, device="cuda"), None, None)
    for _ in range(5): eng.process(d0, d1, NB, out=out)
    eng.sync(); eng.reset_counters(); eng.enable_timing(True)
    for _ in range(20): eng.process(d0, d1, NB, out=out)
    eng.sync()
    ms, n = eng.dominant_kernel_time()
    us = 1e3 * ms / n
    print(f"{NB:4d} block pairs ({2 * NB * 2 * S / 1e6:6.0f} MB of raw bytes): fused kernel {us:7.1f} us/launch  {us / NB * 1e3:7.1f} ns/block  {NB * S / us:8.0f} Msamples/s")
    eng.close()
