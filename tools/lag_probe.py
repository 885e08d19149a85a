#!/usr/bin/env python
"""Lag search timing (BASELINE config 2): 1 block and 92 blocks accumulated, n = 2^18 (2^19-point transforms)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

n, nb = 262144, 92
raw0, raw1 = synth.tiled_recording(nb, n, base_blocks=4, delay=37, seed=99)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
reps = 2 if "--ncu" in sys.argv else 20
for label, kw in (("head/tail kernels", {}), ("generic passes", {"force_generic": True})):
    eng = FxEngine(n, 4096, 1, max_blocks=nb, **kw)
    for blocks in (1, nb):
        r = eng.lag(d0, d1, blocks); eng.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = eng.lag(d0, d1, blocks)
        dt = (time.perf_counter() - t0) / reps
        # device time of the asynchronous half (accumulate only), events on the engine's stream
        x = eng.lag_accumulate(d0, d1, blocks); eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.lag_accumulate(d0, d1, blocks, xacc=x, first=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{label:18s} blocks={blocks:3d}  lag() {dt*1e6:9.1f} us/call (synchronous, lag={r[0]-r[1]})   accumulate {e0.elapsed_time(e1)/reps*1e3:9.1f} us (device)")
    eng.close()
