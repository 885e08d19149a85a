#!/usr/bin/env python
"""Does the byte-sum pre-pass of call k+1 run while the fused kernel of call k is still running?
Events on the library's two streams after every call; prints the timeline (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

S, N, NB = 262144, 4096, 550
raw0, raw1 = synth.tiled_recording(NB, S, base_blocks=4)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=NB)
out = (torch.empty((NB, N), dtype=torch.complex64, device="cuda"), None, None)
for _ in range(3):
    eng.process(d0, d1, NB, out=out, inputs_ready=True)
eng.sync()
K = 8
t0 = torch.cuda.Event(enable_timing=True)
ea = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
em = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
t0.record(eng.stream)
for k in range(K):
    eng.process(d0, d1, NB, out=out, inputs_ready=True)
    ea[k].record(eng.stream_aux)
    em[k].record(eng.stream)
eng.sync()
for k in range(K):
    print(f"call {k}: pre-pass done at {t0.elapsed_time(ea[k])*1e3:8.1f} us   rows done at {t0.elapsed_time(em[k])*1e3:8.1f} us")
