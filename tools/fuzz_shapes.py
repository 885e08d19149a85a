#!/usr/bin/env python
"""Random shapes through the fast paths (fused 256..4096 bins, head/tail 8192..65536) against the unfused
kernels on the same handle parameters: rows and autos must agree to float32 rounding.  Run under gpurun:
    python tools/fuzz_shapes.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from effex_b200.engine import FxEngine

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst = 0.0
for case in range(n_cases):
    N = int(2 ** rng.integers(8, 17))
    P = int(rng.integers(1, 70)) if N <= 16384 else int(rng.integers(1, 9))
    S = P * N + 8 * int(rng.integers(0, N // 8))           # ragged tail, still a multiple of 8
    nb = int(rng.integers(1, 6)) if S * 1 > 2**20 else int(rng.integers(1, 40))
    nb = max(1, min(nb, (1 << 25) // S))
    dc = bool(rng.integers(0, 2))
    raw0 = rng.integers(0, 256, 2 * S * nb, dtype=np.uint8)
    raw1 = rng.integers(0, 256, 2 * S * nb, dtype=np.uint8)
    raw1[: 2 * S * nb - 10] = raw0[10:]                      # correlated pair with a 5-sample lag
    d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
    fast = FxEngine(S, N, 4, max_blocks=nb, dc_remove=dc)
    slow = FxEngine(S, N, 4, max_blocks=nb, dc_remove=dc, force_generic=True)
    for e in (fast, slow):
        e.set_delay(2.4e6, 1.4204e9, 5 / 2.4e6)
    xf, af0, af1 = (v.cpu().numpy() for v in fast.process(d0, d1, nb, autos=True))
    xs, as0, as1 = (v.cpu().numpy() for v in slow.process(d0, d1, nb, autos=True))
    e_x = float(np.abs(xf - xs).max() / np.abs(xs).max())
    e_a = float(max(np.abs(af0 - as0).max() / as0.max(), np.abs(af1 - as1).max() / as1.max()))
    worst = max(worst, e_x, e_a)
    flag = "" if max(e_x, e_a) < 2e-5 else "   <-- MISMATCH"
    print(f"case {case:3d}: N={N:6d} P={P:3d} S={S:9d} blocks={nb:3d} dc={int(dc)}  cross {e_x:.2e} autos {e_a:.2e}{flag}")
    fast.close(); slow.close()
    assert not flag, "fast path disagrees with the unfused kernels"
print(f"fuzz ok: {n_cases} cases, worst relative difference {worst:.2e}")
