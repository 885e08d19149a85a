#!/usr/bin/env python
"""Lag search on the head/tail kernels against the unfused passes for transform lengths covering every split of the
register head (M = G*4096, G = 2 ... 256), ragged n, several blocks, raw bytes and complex input.
usage: python tools/fuzz_lag.py [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fx_oracle as orc
from effex_b200 import synth
from effex_b200.engine import FxEngine

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(seed)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sizes = [4096, 5000, 8192, 12000, 16384, 30000, 32768, 50000, 2**16, 100000, 2**17, 200000, 2**18, 400000, 2**19]
worst = 0.0
for n in sizes:
    nblk = int(rng.integers(1, 4)) if n <= 2**17 else 1
    delay = int(rng.integers(-min(n // 4, 3000), min(n // 4, 3000)))
    raw0, raw1 = synth.correlated_pair(nblk * n, delay=delay, dc0=0.01 - 0.02j, seed=int(rng.integers(1 << 30)))
    fast, slow = FxEngine(n, 8, 1, max_blocks=nblk), FxEngine(n, 8, 1, max_blocks=nblk, force_generic=True)
    a, b = fast.lag(dev(raw0), dev(raw1), nblk), slow.lag(dev(raw0), dev(raw1), nblk)
    assert a[1] == b[1] and a[0] - a[1] == delay, (n, delay, a, b)
    rel = max(abs(x - y) for x, y in zip(a[2:], b[2:])) / max(b[2:])
    worst = max(worst, rel)
    assert rel <= 1e-5, (n, a, b)
    x0 = torch.from_numpy(orc.block_from_u8(raw0[:2 * n]).astype(np.complex64)).cuda()
    x1 = torch.from_numpy(orc.block_from_u8(raw1[:2 * n]).astype(np.complex64)).cuda()
    assert fast.lag(x0, x1)[1] == slow.lag(x0, x1)[1]
    print(f"n={n:7d} M={fast.lag_fft_len():8d} blocks={nblk} delay={delay:6d}  ok  neighbours rel {rel:.2e}")
    fast.close(); slow.close()
print(f"lag fuzz ok: {len(sizes)} sizes, worst neighbour difference {worst:.2e}")
