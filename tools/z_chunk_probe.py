#!/usr/bin/env python
"""Does the intermediate Z of the 8192..65536-bin path stay in L2 when it is walked in small frame chunks?
Streaming span at 65536 bins (run_big_span chunks by frames); EFFEX_FX_Z_ELEMS sets the chunk."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine
S, N, nb = 2**24, 65536, 2
raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=1, seed=5)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=nb)
acc = eng.new_accumulators()
sums = eng.span_sums(d0, d1, nb)
for _ in range(2):
    eng.integrate_stream(d0, d1, acc, nb, sums=sums, total_samp=nb * S)
eng.sync()
eng.reset_counters(); eng.enable_timing(True)
t0 = time.perf_counter()
for _ in range(10):
    eng.integrate_stream(d0, d1, acc, nb, sums=sums, total_samp=nb * S)
eng.sync()
dt = (time.perf_counter() - t0) / 10
ms, n = eng.dominant_kernel_time()
print(f"Z elems {os.environ.get('EFFEX_FX_Z_ELEMS', 'default 2^26')}: {dt*1e6:8.1f} us/pass  {nb*S/dt/1e6:9.0f} Msamples/s   tail kernel total {ms/10*1e3:7.1f} us in {n//10} launches")
