"""The reference's own DSP tests (reference tests/test_effex.py:62-121), pointed at
the drop-in `Correlator`, plus the loop/CSV behaviour around the hot path."""
import numpy as np
import pytest
import torch

from oracle import fx_oracle as orc
from effex_b200 import synth, csvio
from effex_b200.correlator import Correlator

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cor():
    c = Correlator()
    yield c
    c.close()


def test_correlator_init(cor):
    # reference tests/test_effex.py:127-134
    assert cor.state == 'OFF'
    assert cor.mode == 'SPECTRUM'
    assert cor.bandwidth == 2.4e6
    assert cor.nbins == 2**12
    assert cor.frequency == 1.4204e9
    assert cor.gain == 49.6


@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [1e6, 2.4e6])
@pytest.mark.parametrize('freq', [2e4, 1e5])
@pytest.mark.parametrize('taps', [4, 32])
@pytest.mark.parametrize('branches', [2048, 4096])
def test_func_spectrometer_poly(cor, num_samp, rate, freq, taps, branches):
    iq = synth.complex_sinusoid(num_samp, rate, freq)
    window = orc.pfb_window(taps, branches)
    spec = cor._spectrometer_poly(iq, taps, branches, window)
    psd = torch.real(spec * torch.conj(spec)).mean(axis=0).cpu().numpy()
    freqs = np.fft.fftshift(np.fft.fftfreq(len(psd), d=1 / rate))
    psd = np.fft.fftshift(psd)
    freq_err_pct = 100. * abs(freqs[np.argmax(psd)] - freq) / freq
    assert freq_err_pct < 1.
    # and elementwise against the oracle's channelizer
    ref = orc.spectrometer_poly(iq, taps, branches, window)
    got = spec.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [2.4e6])
@pytest.mark.parametrize('samp_offset_int', [-2000, -1001, -1, 0, 1, 999, 2000])
def test_func_estimate_delay_gaussian(cor, num_samp, rate, samp_offset_int):
    iq_0, iq_1 = synth.rolled_pair(num_samp, samp_offset_int)
    est_delay = cor._estimate_delay_gaussian(iq_0, iq_1, rate)
    assert abs(samp_offset_int - est_delay * rate) < 0.5
    assert abs(est_delay - orc.estimate_delay_gaussian(iq_0, iq_1, rate)) * rate < 1e-4     # samples


@pytest.mark.parametrize('num_samp', [3 + 2**12, 2**18])
@pytest.mark.parametrize('rate', [2.4e6])
@pytest.mark.parametrize('samp_offset_int', [-2000, -1001, -1, 0, 1, 999, 2000])
def test_func_estimate_delay(cor, num_samp, rate, samp_offset_int):
    iq_0, iq_1 = synth.rolled_pair(num_samp, samp_offset_int)
    est_delay = cor._estimate_delay(iq_0, iq_1, rate)
    assert abs(samp_offset_int / rate - est_delay) < 1e-6


def test_off_nominal_init():
    # reference tests/test_effex.py:225-248
    with pytest.raises(ValueError):
        Correlator(run_time=0)
    Correlator(bandwidth=3.0e6)                # only warns
    with pytest.raises(ValueError):
        Correlator(mode='FOO')
    assert Correlator(mode='continuum').mode == 'CONTINUUM'
    assert Correlator(num_samp=2**20).num_samp == 2**18          # clamp, effex.py:277-284
    assert Correlator(num_samp=100, nbins=16).num_samp == 2**8
    assert Correlator(num_samp=2**20, extended=True).num_samp == 2**20


def test_run_recording_and_csv(tmp_path):
    """effex spectrum mode end to end: first block calibrates (no row), the rest give
    rows; the .csv reads back the way effex.py:798 reads it and matches the oracle."""
    S, N, nb = 2**16, 4096, 5
    raw0, raw1 = synth.correlated_pair(nb * S, delay=37)
    out = tmp_path / "vis.csv"
    cor = Correlator(run_time=1, num_samp=S, nbins=N, output_file=str(out), batch_blocks=3)
    rows = cor.run_recording(raw0, raw1)
    assert rows.shape == (nb - 1, N)
    assert round(cor.calibrated_delay * cor.bandwidth) == 37
    x0, x1 = orc.block_from_u8(raw0[:2 * S]), orc.block_from_u8(raw1[:2 * S])
    tau_ref = orc.estimate_delay(x0, x1, 2.4e6)
    assert abs(cor.calibrated_delay - tau_ref) * 2.4e6 < 1e-3
    ref = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, cor.calibrated_delay, 4, 1, nb - 1)
    for b in range(nb - 1):
        assert np.abs(rows[b] - ref[b]).max() <= 1e-4 * np.abs(ref[b]).max()
    meta, back = csvio.read_rows(str(out))
    assert meta["mode"] == "SPECTRUM" and int(meta["resolution"]) == N
    np.testing.assert_array_equal(back, rows.astype(np.complex128))
    with open(out) as fh:
        assert fh.readline() + fh.readline() == orc.csv_metadata(1, 2.4e6, 1.4204e9, S, N, 49.6, "SPECTRUM")
    cor.close()


def test_continuum_and_test_modes(tmp_path):
    S, N, nb = 2**14, 1024, 4
    raw0, raw1 = synth.correlated_pair(nb * S, delay=5)
    cor = Correlator(num_samp=S, nbins=N, mode='continuum', output_file=str(tmp_path / "c.csv"))
    vis = cor.run_recording(raw0, raw1)
    assert vis.shape == (nb - 1,)
    for b in range(1, nb):
        ref = orc.process_block_u8(raw0[2 * S * b:2 * S * (b + 1)], raw1[2 * S * b:2 * S * (b + 1)], N, 2.4e6,
                                   1.4204e9, cor.calibrated_delay, mode="CONTINUUM")
        assert abs(vis[b - 1] - ref) <= 1e-4 * abs(ref)
    meta, back = csvio.read_rows(str(tmp_path / "c.csv"))
    assert back.shape == (nb - 1, 1)
    cor.close()
    cor = Correlator(num_samp=S, nbins=N, mode='test', output_file=str(tmp_path / "t.csv"))
    vis = cor.run_recording(raw0, raw1)
    tau0 = orc.estimate_delay(orc.block_from_u8(raw0[:2 * S]), orc.block_from_u8(raw1[:2 * S]), 2.4e6,
                              mode="TEST", frequency=1.4204e9)
    for b in range(1, nb):
        tau = tau0 + b * cor.test_delay_sweep_step
        ref = orc.process_block_u8(raw0[2 * S * b:2 * S * (b + 1)], raw1[2 * S * b:2 * S * (b + 1)], N, 2.4e6,
                                   1.4204e9, tau, mode="TEST")
        assert abs(vis[b - 1] - ref) <= 2e-4 * abs(ref) + 1e-12
    cor.close()


def test_run_files_streams_like_run_recording(tmp_path):
    """File ingest (chunked reader thread -> pinned buffers -> fx_process_host) gives the rows of
    the in-memory run."""
    from effex_b200.correlator import run_files
    S, N, nb = 2**15, 1024, 9
    raw0, raw1 = synth.correlated_pair(nb * S, delay=11, seed=21)
    p0, p1 = tmp_path / "c0.iq", tmp_path / "c1.iq"
    raw0.tofile(p0); raw1.tofile(p1)
    c1 = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "a.csv"), batch_blocks=4)
    rows_mem = c1.run_recording(raw0, raw1)
    c2 = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "b.csv"), batch_blocks=4)
    rows_file = run_files(c2, str(p0), str(p1))
    assert rows_file.shape == (nb - 1, N)
    assert c1.calibrated_delay == c2.calibrated_delay
    np.testing.assert_array_equal(rows_mem, rows_file)
    assert open(tmp_path / "a.csv", 'rb').read() == open(tmp_path / "b.csv", 'rb').read()
    c1.close(); c2.close()


def test_run_files_over_fifos_equals_files(tmp_path):
    """Stream ingest (the reference's `_streaming`, effex.py:630-664): two FIFOs fed piecewise --
    what two `rtl_sdr -d K -` processes look like -- give the rows, the calibrated delay and the CSV bytes of the
    same data read from regular files; the first delivered block calibrates and yields no row (:399-401)."""
    import os
    import threading
    from effex_b200.correlator import run_files
    S, N, nb = 2**15, 1024, 9
    raw0, raw1 = synth.correlated_pair(nb * S, delay=11, seed=21)
    p0, p1 = tmp_path / "c0.iq", tmp_path / "c1.iq"
    raw0.tofile(p0); raw1.tofile(p1)
    cf = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "files.csv"), batch_blocks=4)
    rows_file = run_files(cf, str(p0), str(p1))
    fifos = [str(tmp_path / "c0.fifo"), str(tmp_path / "c1.fifo")]
    for f in fifos:
        os.mkfifo(f)

    def feed(path, data, piece):
        with open(path, 'wb', buffering=0) as fh:
            for i in range(0, len(data), piece):
                fh.write(data[i:i + piece].tobytes())
    ths = [threading.Thread(target=feed, args=(fifos[0], raw0, 50000)),
           threading.Thread(target=feed, args=(fifos[1], raw1, 77777))]
    for t in ths:
        t.start()
    cs = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "fifo.csv"), batch_blocks=4)
    rows_fifo = run_files(cs, fifos[0], fifos[1])
    for t in ths:
        t.join(timeout=20)
    assert rows_fifo.shape == (nb - 1, N) and cs.calibrated_delay == cf.calibrated_delay
    # the stream's first batch loses a block to the calibration, so the batches -- and with them the segment
    # plans and the float32 summation order -- differ from the file run: equal to rounding, not bit for bit
    np.testing.assert_allclose(rows_fifo, rows_file, rtol=0, atol=2e-6 * np.abs(rows_file).max())
    from effex_b200 import csvio
    _, back = csvio.read_rows(str(tmp_path / "fifo.csv"))
    np.testing.assert_array_equal(back, rows_fifo.astype(np.complex128))      # the CSV holds exactly these rows
    cf.close(); cs.close()
