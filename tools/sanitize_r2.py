#!/usr/bin/env python
"""Small shapes of the round-2 kernels for compute-sanitizer, each checked against the oracle / a second path:
  - the auto-power-free fused kernel and finalize_integrate_kernel with its tail (ticket counters, local
    accumulators), cross-only accumulators;
  - the mailbox reduce: two handles on one GPU as two ranks, pipelined epochs (push from the finalize tail,
    deferred rank-ordered fold, back-pressure), the generic push/fold kernels, the 8192-bin staged path;
  - the lag search on the head/tail kernels (raw bytes and complex input, accumulate + finish);
  - a Bluestein (non-power-of-two) shape."""
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

# rows + accumulators through the integrate tail, full and cross-only
for N, S, nb in ((4096, 3 * 4096, 5), (1024, 7 * 1024, 4), (256, 19 * 256 + 8, 3)):
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.01j, seed=N)
    for lean in (False, True):
        eng = FxEngine(S, N, 4, max_blocks=nb, cross_only=lean)
        acc = eng.new_accumulators()
        x = eng.process(dev(raw0), dev(raw1), nb, acc=acc).cpu().numpy()
        eng.integrate(dev(raw0), dev(raw1), acc, nb)
        ref = orc.process_recording_u8(raw0, raw1, S, N, BW, FC, 0.0, 4, 0, nb)
        xi, a0, _ = FxEngine.finish_integration(acc)
        assert close(x, ref) and close(xi, ref.mean(axis=0)), (N, lean)
        assert acc["frames"].item() == 2 * nb * (S // N) and (a0.max() == 0) == lean
        eng.close()
    print(f"integrate tail ok N={N}")

# two ranks on one GPU
for N, S in ((4096, 2 * 4096), (8192, 2 * 8192), (128, 1024)):
    nb = 3
    raws = [synth.correlated_pair(nb * S, delay=2, seed=50 + r) for r in range(2)]
    engs = [FxEngine(S, N, 4, max_blocks=nb) for _ in range(2)]
    toks = [e.comm_export(2) for e in engs]
    for r, e in enumerate(engs):
        e.comm_attach(r, 2, toks)
    dv = [(dev(a), dev(b)) for a, b in raws]
    acc = engs[0].new_accumulators()
    for ep in range(4):
        for r in ((0, 1) if ep % 2 == 0 else (1, 0)):
            engs[r].process_reduce(dv[r][0], dv[r][1], nb, acc=acc if r == 0 else None, root=0)
    bufs = [torch.full((2 * N - 5,), float(r + 1), dtype=torch.float64, device="cuda") for r in range(2)]
    for r in (1, 0):
        engs[r].reduce_inplace(bufs[r], root=0)
    for e in engs:
        e.sync()
    want = None
    for r in range(2):
        part = engs[r].new_accumulators()
        engs[r].integrate(dv[r][0], dv[r][1], part, nb)
        engs[r].sync()
        want = part["flat"] if want is None else want + part["flat"]
    np.testing.assert_allclose(acc["flat"].cpu().numpy(), 4 * want.cpu().numpy(), rtol=1e-12, atol=1e-9)
    assert torch.all(bufs[0] == 3.0)
    for e in engs:
        e.close()
    print(f"mailbox reduce ok N={N}")

# lag search on the head/tail kernels
for n, nblk in ((2**13, 3), (3 + 2**12, 2)):
    raw0, raw1 = synth.correlated_pair(nblk * n, delay=-7, seed=9)
    fast, slow = FxEngine(n, 8, 1, max_blocks=nblk), FxEngine(n, 8, 1, max_blocks=nblk, force_generic=True)
    a, b = fast.lag(dev(raw0), dev(raw1), nblk), slow.lag(dev(raw0), dev(raw1), nblk)
    assert a[1] == b[1] and a[0] - a[1] == -7, (a, b)
    x0 = torch.from_numpy(orc.block_from_u8(raw0[:2 * n]).astype(np.complex64)).cuda()
    x1 = torch.from_numpy(orc.block_from_u8(raw1[:2 * n]).astype(np.complex64)).cuda()
    assert fast.lag(x0, x1)[1] == slow.lag(x0, x1)[1]
    xa = fast.lag_accumulate(dev(raw0), dev(raw1), nblk)
    assert fast.lag_finish(xa)[1] == a[1]
    fast.close(); slow.close()
    print(f"lag head/tail ok n={n}")

# Bluestein
S, N, nb = 1000 * 6, 1000, 2
raw0, raw1 = synth.correlated_pair(nb * S, delay=1, seed=2)
eng = FxEngine(S, N, 4, max_blocks=nb)
x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
assert close(x, orc.process_recording_u8(raw0, raw1, S, N, BW, FC, 0.0, 4, 0, nb))
eng.close()
print("bluestein ok")
print("sanitize_r2 ok")
