"""Pin the CPU oracle (oracle/fx_oracle.py).

(1) the reference's own DSP property tests, re-expressed against the oracle
    (reference tests/test_effex.py:62-121: 32 + 14 + 14 parametrisations);
(2) fixtures produced by executing the reference's unmodified effex.py behind
    numpy stand-ins (tests/golden/make_golden.py);
(3) the one third-party routine on the path, cuSignal's channelize_poly, against a thread-by-thread loop
    form of cuSignal's kernel that shares no code with the oracle (tests/golden/cusignal_standin.py) --
    the same stand-in make_golden.py installs, so (2) is not circular either.
"""
import os

import numpy as np
import pytest

from oracle import fx_oracle as orc
from effex_b200 import synth


# ---- (1) reference property tests -----------------------------------------
@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [1e6, 2.4e6])
@pytest.mark.parametrize('freq', [2e4, 1e5])
@pytest.mark.parametrize('taps', [4, 32])
@pytest.mark.parametrize('branches', [2048, 4096])
def test_func_spectrometer_poly(num_samp, rate, freq, taps, branches):
    # reference tests/test_effex.py:62-84
    iq = synth.complex_sinusoid(num_samp, rate, freq)
    window = orc.pfb_window(taps, branches)
    spec = orc.spectrometer_poly(iq, taps, branches, window)
    assert spec.shape == (num_samp // branches, branches)
    psd = np.real(spec * np.conj(spec)).mean(axis=0)
    freqs = np.fft.fftshift(np.fft.fftfreq(len(psd), d=1 / rate))
    psd = np.fft.fftshift(psd)
    freq_err_pct = 100. * abs(freqs[np.argmax(psd)] - freq) / freq
    assert freq_err_pct < 1.


@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [2.4e6])
@pytest.mark.parametrize('samp_offset_int', [-2000, -1001, -1, 0, 1, 999, 2000])
def test_func_estimate_delay_gaussian(num_samp, rate, samp_offset_int):
    # reference tests/test_effex.py:92-106
    iq_0, iq_1 = synth.rolled_pair(num_samp, samp_offset_int)
    est = orc.estimate_delay_gaussian(iq_0, iq_1, rate)
    assert abs(samp_offset_int - est * rate) < 0.5
    n, imax, *_ = orc.lag_search(iq_0, iq_1)
    assert n - imax == samp_offset_int          # integer lag, exact


@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [2.4e6])
@pytest.mark.parametrize('samp_offset_int', [-2000, -1001, -1, 0, 1, 999, 2000])
def test_func_estimate_delay(num_samp, rate, samp_offset_int):
    # reference tests/test_effex.py:109-121
    iq_0, iq_1 = synth.rolled_pair(num_samp, samp_offset_int)
    est = orc.estimate_delay(iq_0, iq_1, rate)
    assert abs(samp_offset_int / rate - est) < 1e-6


def test_wrong_sign_convention_fails():
    """The tone test discriminates the conj convention: the un-conjugated
    variant puts a +f tone in the -f bin."""
    iq = synth.complex_sinusoid(2**14, 2.4e6, 1e5)
    w = orc.pfb_window(4, 2048)
    spec = np.conj(orc.spectrometer_poly(np.conj(iq), 4, 2048, w))  # opposite convention
    psd = np.fft.fftshift((spec * spec.conj()).real.mean(axis=0))
    freqs = np.fft.fftshift(np.fft.fftfreq(2048, d=1 / 2.4e6))
    assert freqs[np.argmax(psd)] < 0


# ---- (2) fixtures from the executed reference -------------------------------
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_golden_reference_case(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ref_case_{tag}.npz"))
    S, N, bw, fc = int(g["S"]), int(g["N"]), float(g["bw"]), float(g["fc"])
    x0 = orc.block_from_u8(g["raw0"])
    x1 = orc.block_from_u8(g["raw1"])
    w = orc.pfb_window(4, N)
    np.testing.assert_array_equal(w, g["window"])
    np.testing.assert_allclose(orc.spectrometer_poly(x0, 4, N, w), g["spec0"], rtol=0, atol=1e-15)
    x = orc.pfb_xcorr(x0, x1, 4, N, w, bw, fc, 0.0)
    np.testing.assert_allclose(x, g["xspec_tau0"], rtol=1e-13, atol=1e-18)
    tau = orc.estimate_delay(x0, x1, bw)
    assert tau == pytest.approx(float(g["calibrated_delay"]), rel=1e-12, abs=1e-18)
    assert orc.estimate_delay_gaussian(x0, x1, bw) == pytest.approx(float(g["delay_gauss"]), rel=1e-12, abs=1e-18)
    xc = orc.pfb_xcorr(x0, x1, 4, N, w, bw, fc, float(g["calibrated_delay"]))
    np.testing.assert_allclose(xc, g["xspec_cal"], rtol=1e-12, atol=1e-18)
    vis = orc.pfb_xcorr(x0, x1, 4, N, w, bw, fc, float(g["calibrated_delay"]), mode="continuum")
    np.testing.assert_allclose(vis, g["vis_continuum"], rtol=1e-12)
    assert orc.test_delay_offset(fc) == pytest.approx(float(g["test_delay_offset"]), rel=1e-15)
    assert orc.estimate_delay(x0, x1, bw, mode="TEST", frequency=fc) == pytest.approx(
        float(g["delay_test_mode"]), rel=1e-12)
    # whole chain from bytes
    np.testing.assert_allclose(
        orc.process_block_u8(g["raw0"], g["raw1"], N, bw, fc, float(g["calibrated_delay"])),
        g["xspec_cal"], rtol=1e-12, atol=1e-18)


@pytest.mark.parametrize('tag,mode', [('a', 'SPECTRUM'), ('b', 'SPECTRUM'), ('c', 'CONTINUUM')])
def test_golden_csv_bytes(golden_dir, tag, mode):
    with open(os.path.join(golden_dir, f"ref_meta_{tag}.csv")) as fh:
        ref = fh.read()
    if tag == 'c':
        got = orc.csv_metadata(2, 2.4e6, 1.4204e9, 4096, 256, 49.6, mode)
        assert got == ref
        return
    g = np.load(os.path.join(golden_dir, f"ref_case_{tag}.npz"))
    got = orc.csv_metadata(1, float(g["bw"]), float(g["fc"]), int(g["S"]), int(g["N"]), 49.6, mode)
    got += orc.csv_row(g["xspec_cal"])
    assert got == ref
    # and it reads back the way effex.py:798 / post_process.py:219 read it
    import io
    back = np.loadtxt(io.StringIO(ref), dtype=np.complex128, delimiter=',', skiprows=2)
    np.testing.assert_array_equal(back, g["xspec_cal"])


def test_pfb_fir_form_matches_channelizer():
    """F[i,c] = exp(-2j pi c/N) * FFT_p(w[i,:])[c]."""
    rng = np.random.default_rng(1)
    N, T, P = 64, 4, 9
    x = rng.normal(size=N * P + 5) + 1j * rng.normal(size=N * P + 5)
    h = orc.pfb_window(T, N)
    F = orc.spectrometer_poly(x, T, N, h)
    w = orc.pfb_fir(x, h, N)
    c = np.arange(N)
    np.testing.assert_allclose(np.exp(-2j * np.pi * c / N) * np.fft.fft(w, axis=1), F, atol=1e-12)


# ---- (3) the oracle's channelizer vs an independent restatement of cuSignal's kernel --------------------------
@pytest.mark.parametrize("N,T,S", [(64, 4, 1000), (256, 4, 4096), (32, 32, 32 * 40 + 5), (16, 9, 16 * 30),
                                   (1024, 4, 8192), (8, 1, 64), (128, 16, 128 * 3), (64, 4, 64 * 2)])
def test_channelize_poly_equals_the_kernel_shaped_stand_in(N, T, S):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import cusignal_standin
    rng = np.random.default_rng(N + T)
    x = rng.normal(size=S) + 1j * rng.normal(size=S)
    h = orc.pfb_window(T, N) * (1 + 0.1 * rng.normal(size=T * N))      # no symmetry to hide an index reversal behind
    a, b = cusignal_standin.channelize_poly(x, h, N), orc.channelize_poly(x, h, N)
    assert a.shape == b.shape == (N, S // N)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-13 * np.abs(b).max())
    with pytest.raises(NotImplementedError):
        cusignal_standin.channelize_poly(x, np.ones(33 * 8), 8)


def test_golden_fixtures_do_not_come_from_the_oracle_channelizer():
    src = open(os.path.join(os.path.dirname(__file__), "golden", "make_golden.py")).read()
    assert "cusignal_standin.channelize_poly" in src and "= orc.channelize_poly" not in src
