#!/usr/bin/env python
"""Small shapes for compute-sanitizer: the fused kernel at every supported nbins (partial last
super-frame, several segments per CTA) and the generic radix-pass FFT, each checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fx_oracle as orc
from effex_b200 import synth
from effex_b200.engine import FxEngine

worst = 0.0
for S, N, nb in [(3 * 4096, 4096, 3), (5 * 2048, 2048, 3), (7 * 1024, 1024, 3), (11 * 512, 512, 3),
                 (19 * 256 + 8, 256, 3), (2 * 16384, 16384, 1)]:
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.01j, dc1=-0.02, seed=3)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    eng.set_delay(2.4e6, 1.4204e9, 3 / 2.4e6)
    x = eng.process(torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda(), nb).cpu().numpy()
    ref = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, 3 / 2.4e6, 4, 0, nb)
    err = float(np.abs(x - ref).max() / np.abs(ref).max())
    worst = max(worst, err)
    print(f"S={S} N={N} fused={eng.fused} rel err {err:.2e}")
    eng.close()
assert worst < 1e-4
# lag search: forward + inverse radix passes at 2n = 2^14, accumulated over 2 blocks
S = 2**13
raw0, raw1 = synth.correlated_pair(2 * S, delay=5, seed=4)
eng = FxEngine(S, 1024, 4, max_blocks=2)
n, imax, *_ = eng.lag(torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda(), 2)
assert n - imax == 5, (n, imax)
print("lag ok:", n - imax)
eng.close()
print("sanitize_small ok, worst rel err %.2e" % worst)
