#!/usr/bin/env python
"""Small shapes of the kernels added late in round 2 for compute-sanitizer, each checked against the oracle or a
second path:
  - bigfft::head2_kernel (persistent, tensor-memory FIR state, TMA ring with full/empty mbarriers, producer warp)
    + tail_kernel at 8192 / 16384 / 65536 bins, blocks shorter and longer than the 3 warm-up frames, several
    segments per CTA; a streaming span walked in chunks of Z with a halo and recording-wide sums;
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

# streaming span at 8192 bins walked in chunks of Z, second half with its halo and the recording-wide sums
os.environ["EFFEX_FX_Z_ELEMS"] = str(16 * 8192)
S, N, nb = 10 * 8192, 8192, 4
raw0, raw1 = synth.correlated_pair(nb * S, delay=7, dc0=0.02 + 0.01j, seed=23)
d0, d1 = dev(raw0), dev(raw1)
eng = FxEngine(S, N, 4, max_blocks=nb)
w = orc.pfb_window(4, N)
f0, f1 = (orc.spectrometer_poly(orc.block_from_u8(r), 4, N, w) for r in (raw0, raw1))
ref = np.fft.fftshift((f0 * np.conj(f1)).mean(axis=0))
sums = eng.span_sums(d0, d1, nb)
tot = eng.new_accumulators()
half, hb = nb // 2, 2 * 3 * N
lo = 2 * S * half
eng.integrate_stream(d0[:lo], d1[:lo], tot, half, sums=sums, total_samp=nb * S)
eng.integrate_stream(d0[lo:], d1[lo:], tot, nb - half, halo0=d0[lo - hb:lo], halo1=d1[lo - hb:lo], sums=sums, total_samp=nb * S)
xt, _, _ = FxEngine.finish_integration(tot)
assert close(xt, ref)
eng.close()
del os.environ["EFFEX_FX_Z_ELEMS"]
print("head2 streaming span ok")

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
