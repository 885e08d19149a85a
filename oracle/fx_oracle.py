"""CPU oracle for the effex FX-correlator hot path.  TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy/scipy *restatement* of the arithmetic on the
reference's hot path (evanmayer/effex, `effex/effex.py`).  It exists so that
the CUDA path can be checked against something; it is never the product:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` leg may import it;
  * nothing under `effex_b200/` imports it, and the product path raises if the
    CUDA library is missing rather than falling back to this code.

Pinning.  The reference ships no golden vectors (`.gitignore:1` ignores *.csv)
and cannot be imported as-is (cupy / cusignal / pyrtlsdr are absent here).
The oracle is pinned two ways (see `tests/golden/make_golden.py` and
`tests/test_oracle_*.py`):

  1. the reference's OWN source file is executed in this container behind
     numpy-backed stand-ins for cupy/cusignal/rtlsdr, and every function below
     that restates a line range of `effex.py` is compared with what that
     unmodified code returns (fixtures committed under `tests/golden/`);
  2. the one third-party routine whose arithmetic lives outside the reference
     (cuSignal `channelize_poly`, version unpinned -- `requirements.txt` is
     empty) is restated from its published algorithm and pinned by the
     reference's own property tests (`tests/test_effex.py:62-121`; 32 + 14 + 14
     parametrisations), which discriminate the sign/ordering conventions, and
     by `tests/golden/cusignal_standin.py`, a thread-by-thread loop form of
     cuSignal's kernel that shares no code with this module; that stand-in --
     not this module -- is what the fixtures of (1) were generated with.

Each function cites the reference `file:line` it follows (paths relative to
the reference checkout).
"""
from __future__ import annotations

import io
import numpy as np
import scipy.fft
import scipy.signal

NTAPS_DEFAULT = 4          # effex/effex.py:115
MODES = ("SPECTRUM", "CONTINUUM", "TEST")   # effex/effex.py:35


# --------------------------------------------------------------------------
# a0: uint8 interleaved IQ -> complex   (pyrtlsdr packed_bytes_to_iq, reached
#     from effex/effex.py:652 `sdr.stream(format='samples', ...)`)
# --------------------------------------------------------------------------
def unpack_iq(raw: np.ndarray) -> np.ndarray:
    """`iq = bytes[0::2] + 1j*bytes[1::2]; iq /= 127.5; iq -= (1+1j)`
    (pyrtlsdr `rtlsdr.py` `packed_bytes_to_iq`, the author's fork is unpinned,
    `install_instructions.md:37-43`)."""
    raw = np.asarray(raw, dtype=np.uint8)
    iq = raw[0::2].astype(np.float64) + 1j * raw[1::2].astype(np.float64)
    iq /= 127.5
    iq -= (1 + 1j)
    return iq


# --------------------------------------------------------------------------
# a1: DC removal                                     effex/effex.py:394-395
# --------------------------------------------------------------------------
def remove_dc(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.complex128)
    return (x.real - x.real.mean()) + 1j * (x.imag - x.imag.mean())


# --------------------------------------------------------------------------
# a2: PFB prototype filter                           effex/effex.py:126-127
#     (same expression in tests/test_effex.py:73-74)
# --------------------------------------------------------------------------
def pfb_window(ntaps: int, nbins: int) -> np.ndarray:
    """`get_window("hamming", T*N) * firwin(T*N, cutoff=1/N, window='rectangular')`.
    cusignal.get_window / cusignal.firwin are GPU ports of the scipy routines
    of the same name and signature (periodic Hamming by default)."""
    L = int(ntaps) * int(nbins)
    return (scipy.signal.get_window("hamming", L)
            * scipy.signal.firwin(L, cutoff=1.0 / nbins, window="rectangular"))


# --------------------------------------------------------------------------
# a3: polyphase channelizer
# --------------------------------------------------------------------------
def channelize_poly(x: np.ndarray, h: np.ndarray, n_chans: int) -> np.ndarray:
    """cuSignal `cusignal.filtering.channelize_poly(x, h, n_chans)` (called at
    effex/effex.py:553).  Published algorithm (cuSignal `_channelizer.cu` +
    `filtering/channelize_poly`): per output frame `i` and branch `m`

        v[i, m] = sum_k conj(h[k*N + m]) * conj(x[(i-k)*N + (N-1-m)])   (x[<0] = 0)

    followed by `conj(fft(v, axis=-1)).T`, i.e. shape (n_chans, n_pts).
    Filters longer than 32 taps per branch are rejected by cuSignal."""
    x = np.asarray(x)
    h = np.asarray(h)
    N = int(n_chans)
    n_taps = int(len(h) / N)
    if n_taps > 32:
        raise NotImplementedError(
            "The number of calculated taps ({}) in each filter is currently "
            "capped at 32".format(n_taps))
    n_pts = int(len(x) / N)
    xr = np.conj(x[: n_pts * N].reshape(n_pts, N)[:, ::-1]).astype(np.complex128)
    hk = np.conj(h[: n_taps * N].reshape(n_taps, N)).astype(np.complex128)
    v = np.zeros((n_pts, N), dtype=np.complex128)
    for k in range(min(n_taps, n_pts)):
        v[k:] += hk[k][None, :] * xr[: n_pts - k]
    return np.conj(scipy.fft.fft(v, axis=-1)).T


def spectrometer_poly(x, ntaps, n_branches, window) -> np.ndarray:
    """effex/effex.py:530-555.  The pad at :551 allocates len+len%N zeros,
    slices back to len(x) and adds x: a no-op.  Returns (P, N)."""
    x = np.asarray(x, dtype=np.complex128)
    x = np.zeros(len(x) + len(x) % n_branches, dtype=np.complex128)[: len(x)] + x
    return channelize_poly(x, window, n_branches).T


def pfb_fir(x: np.ndarray, window: np.ndarray, nbins: int) -> np.ndarray:
    """Pre-FFT polyphase FIR in the GPU-friendly form:
    w[i, p] = sum_k h[k*N + N-1-p] * x[(i-k)*N + p], so that
    spectrometer_poly(x)[i, c] == exp(-2j*pi*c/N) * fft(w[i, :])[c].
    Not a reference function; used to check the kernel's intermediate."""
    x = np.asarray(x, dtype=np.complex128)
    N = int(nbins)
    T = len(window) // N
    P = len(x) // N
    xf = x[: P * N].reshape(P, N)
    hk = np.asarray(window, dtype=np.float64)[: T * N].reshape(T, N)[:, ::-1]
    w = np.zeros((P, N), dtype=np.complex128)
    for k in range(min(T, P)):
        w[k:] += hk[k][None, :] * xf[: P - k]
    return w


# --------------------------------------------------------------------------
# a4: X-engine                                       effex/effex.py:497-527
# --------------------------------------------------------------------------
def rot_vector(nbins, bandwidth, frequency, calibrated_delay) -> np.ndarray:
    """effex/effex.py:516,519 (natural FFT bin order)."""
    freqs = np.fft.fftfreq(nbins, d=1 / bandwidth) + frequency
    return np.exp(-2j * np.pi * freqs * (-calibrated_delay))


def pfb_xcorr(iq0, iq1, ntaps, nbins, window, bandwidth, frequency,
              calibrated_delay=0.0, mode="SPECTRUM"):
    """effex/effex.py:497-527 on two DC-removed complex blocks."""
    f0 = spectrometer_poly(iq0, ntaps, nbins, window)
    f1 = spectrometer_poly(iq1, ntaps, nbins, window)
    rot = rot_vector(f0.shape[-1], bandwidth, frequency, calibrated_delay)
    xpower_spec = f0 * np.conj(f1 * rot)
    xpower_spec = np.fft.fftshift(xpower_spec.mean(axis=0))
    if mode.upper() in ("CONTINUUM", "TEST"):
        return xpower_spec.mean(axis=0) / bandwidth
    return xpower_spec


def auto_powers(iq0, iq1, ntaps, nbins, window):
    """Per-channel auto-power, same normalisation and bin order as pfb_xcorr
    (mean over frames, fftshifted).  Not in the reference (north_star piece c);
    defined by analogy with effex.py:520-521 with f1 := f0."""
    f0 = spectrometer_poly(iq0, ntaps, nbins, window)
    f1 = spectrometer_poly(iq1, ntaps, nbins, window)
    a0 = np.fft.fftshift((f0 * np.conj(f0)).real.mean(axis=0))
    a1 = np.fft.fftshift((f1 * np.conj(f1)).real.mean(axis=0))
    return a0, a1


# --------------------------------------------------------------------------
# a6/a7: delay calibration                           effex/effex.py:558-627
# --------------------------------------------------------------------------
def lag_search(iq0, iq1):
    """effex/effex.py:600-622.  Returns (n, imax, xprev, xbest, xnext)."""
    assert len(iq0) == len(iq1), ("Algorithm assumes input complex timeseries"
                                  " are of equal length.")
    n = len(iq0)
    a = np.zeros(2 * n, dtype=np.complex128)
    b = np.zeros(2 * n, dtype=np.complex128)
    a[0:n] += np.asarray(iq0)
    b[0:n] += np.asarray(iq1)
    f0 = scipy.fft.fft(a)
    f1 = scipy.fft.fft(b)
    xcorr = np.fft.fftshift(scipy.fft.ifft(f0 * np.conj(f1)))
    imax = int(np.argmax(np.abs(xcorr)))
    xprev = np.abs(xcorr[imax - 1])
    xbest = np.abs(xcorr[imax])
    xnext = np.abs(xcorr[imax + 1])        # no bounds handling (TODO at :619)
    return n, imax, xprev, xbest, xnext


def gaussian_peak(n, imax, xprev, xbest, xnext, rate) -> float:
    """effex/effex.py:623-627."""
    delta = 0.5 * (np.log(xprev) - np.log(xnext)) / (
        np.log(xprev) - 2.0 * np.log(xbest) + np.log(xnext))
    return (n - (imax + delta)) / rate


def estimate_delay_gaussian(iq0, iq1, rate) -> float:
    """effex/effex.py:583-627."""
    return gaussian_peak(*lag_search(iq0, iq1), rate)


def test_delay_offset(frequency) -> float:
    """effex/effex.py:151-155."""
    return (1.0 / frequency) / 2 * 1600


test_delay_offset.__test__ = False   # not a pytest test


def estimate_delay(iq0, iq1, rate, mode="SPECTRUM", frequency=1.4204e9) -> float:
    """effex/effex.py:558-580."""
    d = estimate_delay_gaussian(iq0, iq1, rate)
    if mode.upper() == "TEST":
        d -= test_delay_offset(frequency)
    return d


def accumulated_lag_search(blocks0, blocks1):
    """BASELINE config 2: accumulate the 2n-point cross-spectrum over several
    blocks, then ONE inverse FFT and the same argmax/neighbourhood as
    effex.py:613-622.  With one block this is lag_search()."""
    n = len(blocks0[0])
    acc = np.zeros(2 * n, dtype=np.complex128)
    for a0, b0 in zip(blocks0, blocks1):
        a = np.zeros(2 * n, dtype=np.complex128); a[:n] = a0
        b = np.zeros(2 * n, dtype=np.complex128); b[:n] = b0
        acc += scipy.fft.fft(a) * np.conj(scipy.fft.fft(b))
    xcorr = np.fft.fftshift(scipy.fft.ifft(acc))
    imax = int(np.argmax(np.abs(xcorr)))
    return n, imax, np.abs(xcorr[imax - 1]), np.abs(xcorr[imax]), np.abs(xcorr[imax + 1])


# --------------------------------------------------------------------------
# whole-block chain from raw bytes (what one trip of the loop at
# effex/effex.py:388-410 does to a pair of dequeued blocks)
# --------------------------------------------------------------------------
def block_from_u8(raw: np.ndarray) -> np.ndarray:
    return remove_dc(unpack_iq(raw))


def process_block_u8(raw0, raw1, nbins, bandwidth, frequency, calibrated_delay=0.0,
                     mode="SPECTRUM", ntaps=NTAPS_DEFAULT, window=None):
    if window is None:
        window = pfb_window(ntaps, nbins)
    return pfb_xcorr(block_from_u8(raw0), block_from_u8(raw1), ntaps, nbins, window,
                     bandwidth, frequency, calibrated_delay, mode)


def process_recording_u8(raw0, raw1, num_samp, nbins, bandwidth, frequency,
                         calibrated_delay=0.0, ntaps=NTAPS_DEFAULT, first=0, count=None):
    """Rows for consecutive blocks of a raw recording (uint8[2*num_samp*nblocks])."""
    window = pfb_window(ntaps, nbins)
    nblocks = len(raw0) // (2 * num_samp)
    if count is None:
        count = nblocks - first
    rows = []
    for b in range(first, first + count):
        sl = slice(2 * num_samp * b, 2 * num_samp * (b + 1))
        rows.append(process_block_u8(raw0[sl], raw1[sl], nbins, bandwidth, frequency,
                                     calibrated_delay, "SPECTRUM", ntaps, window))
    return np.array(rows)


# --------------------------------------------------------------------------
# CSV                                                effex/effex.py:667-693
# --------------------------------------------------------------------------
def csv_metadata(run_time, bandwidth, frequency, num_samp, nbins, gain, mode) -> str:
    """effex/effex.py:672-684 (header line + frequency-label line)."""
    buf = io.StringIO()
    buf.write((f'run_time:{run_time},'
               + f'bandwidth:{bandwidth},'
               + f'frequency:{frequency},'
               + f'num_samp:{num_samp},'
               + f'resolution:{nbins},'
               + f'gain:{gain},'
               + f'mode:{mode}\n'))
    if 'SPECTRUM' == mode:
        freqs = np.fft.fftshift(np.fft.fftfreq(nbins, d=1 / bandwidth)) + frequency
        np.savetxt(buf, [freqs], delimiter=',')
    else:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np.savetxt(buf, [])
    return buf.getvalue()


def csv_row(data) -> str:
    """effex/effex.py:693 `np.savetxt(fh, [asnumpy(data)], delimiter=',')`."""
    buf = io.StringIO()
    np.savetxt(buf, [np.asarray(data, dtype=np.complex128)], delimiter=',')
    return buf.getvalue()
