"""--resolution that is not a power of two (the reference hands any nbins to cuFFT, effex.py:553, :734):
unfused kernels + a Bluestein transform (chirp-z through power-of-two transforms), against the oracle."""
import numpy as np
import pytest
import torch

from oracle import fx_oracle as orc
from effex_b200 import synth
from effex_b200.engine import FxEngine, pfb_window

pytestmark = pytest.mark.gpu
TOL = 1e-4
BW, FC = 2.4e6, 1.4204e9


def close(got, ref, tol=TOL):
    got = np.asarray(got, dtype=np.complex128); ref = np.asarray(ref, dtype=np.complex128)
    return (np.abs(got - ref).max() <= tol * np.abs(ref).max()
            and np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref))


@pytest.mark.parametrize("nbins,ntaps,S,nb", [(1000, 4, 16000, 3), (3000, 4, 3000 * 9 + 17, 2), (12, 4, 600, 4),
                                              (4095, 4, 4095 * 8, 2), (6000, 5, 6000 * 7, 2), (100, 32, 100 * 40, 2),
                                              (40000, 4, 40000 * 5, 1), (768, 4, 768 * 20, 3)])
def test_any_nbins_rows_and_accumulators(nbins, ntaps, S, nb):
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.02 - 0.01j, dc1=0.015j, seed=nbins)
    tau = 3 / BW
    eng = FxEngine(S, nbins, ntaps, max_blocks=nb)
    assert not eng.fused
    eng.set_delay(BW, FC, tau)
    d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
    acc = eng.new_accumulators()
    x, a0, a1 = eng.process(d0, d1, nb, autos=True, acc=acc)
    x, a0 = x.cpu().numpy(), a0.cpu().numpy()
    w = pfb_window(ntaps, nbins)
    tot = np.zeros(nbins, dtype=np.complex128)
    for b in range(nb):
        sl = slice(2 * S * b, 2 * S * (b + 1))
        assert close(x[b], orc.process_block_u8(raw0[sl], raw1[sl], nbins, BW, FC, tau, "SPECTRUM", ntaps, w)), b
        f0 = orc.spectrometer_poly(orc.block_from_u8(raw0[sl]), ntaps, nbins, w)
        f1 = orc.spectrometer_poly(orc.block_from_u8(raw1[sl]), ntaps, nbins, w)
        assert close(a0[b], np.fft.fftshift((abs(f0) ** 2).mean(axis=0))), b
        tot += (f0 * np.conj(f1)).sum(axis=0)
    xi, _, _ = FxEngine.finish_integration(acc)
    assert acc["frames"].item() == nb * (S // nbins)
    assert close(xi, np.fft.fftshift(tot / (nb * (S // nbins))))
    # the channelizer on its own (what the reference's tone test calls), phase factor included
    sig = orc.block_from_u8(raw0[:2 * S])
    f = eng.pfb(sig).cpu().numpy()
    assert close(f, orc.spectrometer_poly(sig, ntaps, nbins, w), 2e-5)
    eng.close()


def test_correlator_runs_a_recording_at_a_non_power_of_two_resolution(tmp_path):
    from effex_b200.correlator import Correlator
    from effex_b200 import csvio
    S, N, nb = 30000, 1000, 5
    raw0, raw1 = synth.correlated_pair(nb * S, delay=9, seed=3)
    cor = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "v.csv"), batch_blocks=2)
    rows = cor.run_recording(raw0, raw1)
    assert rows.shape == (nb - 1, N) and abs(cor.calibrated_delay * 2.4e6 - 9) < 0.5
    ref = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, cor.calibrated_delay, 4, 1, nb - 1)
    for b in range(nb - 1):
        assert close(rows[b], ref[b])
    meta, back = csvio.read_rows(str(tmp_path / "v.csv"))
    assert meta["resolution"] == "1000"
    np.testing.assert_array_equal(back, rows.astype(np.complex128))
    cor.close()
