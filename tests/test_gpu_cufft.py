"""cuFFT as an ORACLE for the hand-written transforms (north_star (b): "checked against cuFFT only as an
oracle"; the reference calls cuFFT at effex.py:553 through cuSignal and at :611-613 through
cupy).  torch.fft on a CUDA tensor is cuFFT; it appears in tests and tools only, never in the product."""
import numpy as np
import pytest
import torch

from effex_b200 import synth
from effex_b200.engine import FxEngine, pfb_window

pytestmark = pytest.mark.gpu


def _fir_frames(x, ntaps, nbins, h):
    """w[i, p] = sum_k h[kN + N-1-p] x[(i-k)N + p], float64 on the host"""
    P = len(x) // nbins
    xf = x[:P * nbins].reshape(P, nbins)
    w = np.zeros((P, nbins), dtype=np.complex128)
    for k in range(ntaps):
        taps = h[k * nbins:(k + 1) * nbins][::-1]
        w[k:] += taps * xf[:P - k]
    return w


@pytest.mark.parametrize("nbins,ntaps,S", [(4096, 4, 2**16), (2048, 4, 2**15), (1024, 4, 2**14), (256, 8, 2**12),
                                           (8192, 4, 2**17), (65536, 4, 2**19)])
def test_channelizer_fft_equals_cufft(nbins, ntaps, S):
    rng = np.random.default_rng(4)
    x = (rng.normal(size=S) + 1j * rng.normal(size=S)) * 0.3
    h = pfb_window(ntaps, nbins)
    eng = FxEngine(S, nbins, ntaps)
    got = eng.pfb(x)                                                        # hand-written FIR + FFT (+ phase)
    w = torch.from_numpy(_fir_frames(x, ntaps, nbins, h).astype(np.complex64)).cuda()
    c = torch.arange(nbins, device="cuda", dtype=torch.float64)
    phase = torch.exp(-2j * np.pi * c / nbins).to(torch.complex64)
    ref = torch.fft.fft(w, dim=1) * phase                                   # cuFFT C2C, batched
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err
    eng.close()


@pytest.mark.parametrize("n,nblk", [(2**18, 2), (2**14, 3), (3 + 2**12, 2)])
def test_lag_cross_spectrum_equals_cufft(n, nblk):
    """the accumulated 2n-point cross-spectrum of the lag search (fx_lag.cuh) against cuFFT transforms"""
    raw0, raw1 = synth.correlated_pair(nblk * n, delay=21, seed=8)
    d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
    eng = FxEngine(n, 8, 1, max_blocks=nblk)
    xacc = eng.lag_accumulate(d0, d1, nblk)
    M = eng.lag_fft_len()

    def chan(d):
        b = d.view(nblk, n, 2).to(torch.float32)
        b = (b - b.mean(dim=1, keepdim=True)) / 127.5
        return torch.complex(b[..., 0], b[..., 1])
    A, B = torch.fft.fft(chan(d0), n=M, dim=1), torch.fft.fft(chan(d1), n=M, dim=1)      # cuFFT, zero-padded
    ref = (A * B.conj()).sum(dim=0)
    err = float((xacc - ref).abs().max() / ref.abs().max())
    assert err <= 2e-5, err
    # and the whole search against cuFFT's inverse
    xc = torch.fft.fftshift(torch.fft.ifft(ref))
    j0 = M // 2 - n                                              # xc_shift[j] = lag j - n lives at M/2 + (j - n)
    imax_ref = int(torch.argmax(xc.abs()[j0:j0 + 2 * n]).item())
    assert eng.lag_finish(xacc)[1] == imax_ref
    eng.close()
