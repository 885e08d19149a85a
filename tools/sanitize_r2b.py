#!/usr/bin/env python
"""Small shapes of the kernels added late in round 2 for compute-sanitizer, each checked against the oracle or a
second path:
  - bigfft::head2_kernel (persistent, tensor-memory FIR state, TMA ring with full/empty mbarriers, producer warp)
    + tail_kernel at 8192 / 16384 / 65536 bins, blocks shorter and longer than the 3 warm-up frames, several
    segments per CTA;
  - lag::lag_head2_kernel with and without the shared-memory exchange (G = 2, 4, 32, 64), raw bytes and complex
    input, against the unfused passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fx_oracle as orc
from effex_b200 import synth
from effex_b200.engine import FxEngine

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
BW, FC = 2.4e6, 1.4204e9

def close(a, b, tol=1e-4):
    a = np.asarray(a, dtype=np.complex128); b = np.asarray(b, dtype=np.complex128)
    return np.abs(a - b).max() <= tol * np.abs(b).max()

for N, P, nb in ((8192, 5, 3), (16384, 2, 2), (65536, 6, 2)):
    S = P * N
    raw0, raw1 = synth.correlated_pair(nb * S, delay=5, dc0=0.02 - 0.01j, seed=N)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    assert not eng.fused
    x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    ref = orc.process_recording_u8(raw0, raw1, S, N, BW, FC, 0.0, 4, 0, nb)
    assert close(x, ref), N
    eng.close()
    print(f"head2 + tail ok N={N}")

for n, nblk in ((4096, 2), (8192, 3), (2**16, 2), (2**17 - 8, 1)):
    raw0, raw1 = synth.correlated_pair(nblk * n, delay=-7, seed=9)
    fast, slow = FxEngine(n, 8, 1, max_blocks=nblk), FxEngine(n, 8, 1, max_blocks=nblk, force_generic=True)
    a, b = fast.lag(dev(raw0), dev(raw1), nblk), slow.lag(dev(raw0), dev(raw1), nblk)
    assert a[1] == b[1] and a[0] - a[1] == -7, (a, b)
    x0 = torch.from_numpy(orc.block_from_u8(raw0[:2 * n]).astype(np.complex64)).cuda()
    x1 = torch.from_numpy(orc.block_from_u8(raw1[:2 * n]).astype(np.complex64)).cuda()
    assert fast.lag(x0, x1)[1] == slow.lag(x0, x1)[1]
    fast.close(); slow.close()
    print(f"lag_head2 ok n={n}")
