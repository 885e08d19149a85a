#!/usr/bin/env python
"""Latency of ONE block pair through the host entry point (the reference's real-time loop correlates one
block per trip): pinned host bytes in -> row on the host, synchronous.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

for S, N in [(262144, 4096), (262144, 1024), (2**18, 65536)]:
    raw0, raw1 = synth.correlated_pair(S, delay=37)
    h0, h1 = torch.from_numpy(raw0).pin_memory().numpy(), torch.from_numpy(raw1).pin_memory().numpy()
    out = torch.empty((1, N), dtype=torch.complex64).pin_memory().numpy()
    eng = FxEngine(S, N, 4, max_blocks=1)
    eng.set_delay(2.4e6, 1.4204e9, 37 / 2.4e6)
    for _ in range(20):
        eng.process_host(h0, h1, 1, out=out)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        eng.process_host(h0, h1, 1, out=out)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"S={S} N={N}: one block pair host->host  median {np.median(ts):7.1f} us  p99 {np.percentile(ts, 99):7.1f} us"
          f"  ({S / 2.4e6 * 1e3:.0f} ms of signal at 2.4 MS/s)")
    eng.close()
