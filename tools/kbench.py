#!/usr/bin/env python
"""Kernel-time A/B of library variants on the canonical workload (run under gpurun).
usage: python tools/kbench.py variant.so [variant2.so ...]   (each in a fresh process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
from effex_b200 import synth
from effex_b200.engine import FxEngine
S, N, NB = 262144, 4096, 550
raw0, raw1 = synth.tiled_recording(NB, S, base_blocks=4)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=NB)
out = (torch.empty((NB, N), dtype=torch.complex64, device="cuda"), None, None)
for _ in range(3): eng.process(d0, d1, NB, out=out)
eng.sync(); eng.reset_counters(); eng.enable_timing(True)
import time
t0 = time.perf_counter()
for _ in range(10): eng.process(d0, d1, NB, out=out)
eng.sync()
step_us = (time.perf_counter() - t0) / 10 * 1e6
ms, n = eng.dominant_kernel_time()
chk = float(out[0].abs().sum().item())
print("%%-28s fused kernel %%7.1f us/launch %%8.0f Msamples/s (kernel only)   step %%7.1f us %%8.0f Msamples/s  checksum %%.6e" %% (os.path.basename(os.environ.get("EFFEX_FX_LIB","default")), 1e3*ms/n, NB*S/(ms/n*1e-3)/1e6, step_us, NB*S/step_us, chk))
''' % ROOT

for lib in sys.argv[1:]:
    env = dict(os.environ, EFFEX_FX_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
