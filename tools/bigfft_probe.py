#!/usr/bin/env python
"""C3 (nbins 65536, S = 2^24, 2 blocks) through the head/tail kernels -- the command profiled for
profiles/r01_bigfft_ncu.txt:  ncu --set full --clock-control none -k regex:"head_kernel|tail_kernel" -s 2 -c 2 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine
S, N, nb = 2**24, 65536, 2
raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=2)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=nb)
out = (torch.empty((nb, N), dtype=torch.complex64, device="cuda"), None, None)
for _ in range(2):
    eng.process(d0, d1, nb, out=out)
eng.sync()
