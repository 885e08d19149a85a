#!/usr/bin/env python
"""Device-resident throughput of the BASELINE configs and of every kernel family (fused 256..4096 bins,
head/tail 8192..65536 bins, lag search) -- informational."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

def run(name, S, N, nb, reps=5):
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=min(nb, 4))
    d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
    eng = FxEngine(S, N, 4, max_blocks=nb)
    out = (torch.empty((nb, N), dtype=torch.complex64, device="cuda"), None, None)
    eng.process(d0, d1, nb, out=out); eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.process(d0, d1, nb, out=out)
    eng.sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:28s} S={S:9d} N={N:6d} blocks={nb:5d} fused={eng.fused!s:5s} {dt*1e3:9.3f} ms/pass  {nb*S/dt/1e6:10.0f} Msamples/s")
    eng.close()

run("C1 canonical", 262144, 4096, 550)
run("C5 short integrations", 319488, 1024, 600)
run("N=1024 S=2^18", 262144, 1024, 550)
run("N=2048", 262144, 2048, 550)
run("N=512", 262144, 512, 550)
run("N=256", 262144, 256, 550)
run("N=8192", 262144, 8192, 550)
run("N=16384", 2**20, 16384, 140)
run("C3 hi-res line", 2**24, 65536, 2)
n = 262144
eng = FxEngine(n, 4096, 4, max_blocks=92)
raw0, raw1 = synth.tiled_recording(92, n, base_blocks=4)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng.lag(d0, d1, 1); eng.lag(d0, d1, 92)          # warm-up: workspaces of both shapes
t0 = time.perf_counter(); eng.lag(d0, d1, 1); t1 = time.perf_counter(); eng.lag(d0, d1, 92); t2 = time.perf_counter()
print(f"lag search 2n=2^19: 1 block {1e3*(t1-t0):.2f} ms ; C2 (92 blocks accumulated) {1e3*(t2-t1):.1f} ms")
