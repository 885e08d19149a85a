"""BASELINE.json configs at their full sizes, through the C ABI, against the oracle
on what the oracle can finish in seconds and through size-independent properties for the rest."""
import numpy as np
import pytest
import torch

from oracle import fx_oracle as orc
from effex_b200 import synth, sharding
from effex_b200.engine import FxEngine, rot_vector

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, ref, tol=TOL):
    got = np.asarray(got, dtype=np.complex128); ref = np.asarray(ref, dtype=np.complex128)
    return (np.abs(got - ref).max() <= tol * np.abs(ref).max()
            and np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref))


def test_c1_sixty_seconds_of_spectrum_mode():
    """configs[0]: 550 blocks of S=262144, N=4096.  First and last 3 blocks vs the oracle; all rows
    finite; the tiled input (period 8 blocks) must give rows that repeat with that period."""
    S, N, nb = 262144, 4096, 550
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=8, delay=37)
    tau = 37 / 2.4e6
    eng = FxEngine(S, N, 4, max_blocks=nb)
    eng.set_delay(2.4e6, 1.4204e9, tau)
    x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    assert np.isfinite(x.view(np.float32)).all()
    for b in (0, 1, 2, 547, 548, 549):
        sl = slice(2 * S * b, 2 * S * (b + 1))
        assert close(x[b], orc.process_block_u8(raw0[sl], raw1[sl], N, 2.4e6, 1.4204e9, tau)), b
    for b in range(8, nb):
        assert close(x[b], x[b - 8], 2e-6), b
    # the calibrated phase is flat: residual phase = -2 pi fc tau (mod 2 pi) across the band
    ph = np.angle(x[5][N // 4: 3 * N // 4] * np.exp(2j * np.pi * np.mod(1.4204e9 * tau, 1.0)))
    assert np.abs(ph).mean() < 0.2
    eng.close()


def test_c2_ten_second_lag_search():
    """configs[1]: cross-spectrum accumulated over 10 s (92 blocks), one inverse FFT, lag = +37 exactly."""
    S, nb = 262144, 92
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=4, delay=37, seed=99)
    eng = FxEngine(S, 4096, 4, max_blocks=nb)
    n, imax, p, q, r = eng.lag(dev(raw0), dev(raw1), nb)
    assert n - imax == 37
    assert q > 5 * max(p, r)
    # 4 distinct blocks accumulated 23 times = 23 x the 4-block accumulation (linearity)
    n4, imax4, p4, q4, r4 = eng.lag(dev(raw0[:2 * S * 4]), dev(raw1[:2 * S * 4]), 4)
    assert imax4 == imax
    assert q == pytest.approx(23 * q4, rel=1e-4)
    eng.close()


def test_c3_high_resolution_line():
    """configs[2]: resolution 65536, S = 2^24 (extended mode), 4-tap PFB: line at baseband +5752 Hz ->
    natural bin 157; whole row vs the oracle."""
    S, N = 2**24, 65536
    raw0, raw1 = synth.hi_line_pair(S)
    eng = FxEngine(S, N, 4, max_blocks=1)
    assert not eng.fused
    x = eng.process(dev(raw0), dev(raw1), 1).cpu().numpy()[0]
    ref = orc.process_block_u8(raw0, raw1, N, 2.4e6, 1.4204e9, 0.0)
    assert close(x, ref)
    assert int(np.argmax(np.abs(np.fft.ifftshift(x)))) == round(5752.0 / 2.4e6 * N) == 157
    eng.close()


def test_c4_time_sharding_is_invisible():
    """configs[3] (scaled): a long recording cut into contiguous block ranges (as ranks would take them)
    gives the same rows and the same integrated spectrum as one pass."""
    S, N, nb = 262144, 4096, 40
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=5, delay=37, seed=3)
    d0, d1 = dev(raw0), dev(raw1)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    acc1 = eng.new_accumulators()
    rows1 = eng.process(d0, d1, nb, acc=acc1).cpu().numpy()
    x1, a01, a11 = FxEngine.finish_integration(acc1)
    for world in (2, 3, 8):
        acc = eng.new_accumulators()
        rows = []
        for rank in range(world):
            start, count = sharding.shard_range(nb, world, rank)
            part = eng.new_accumulators()
            rows.append(eng.process(d0[2 * S * start:2 * S * (start + count)],
                                    d1[2 * S * start:2 * S * (start + count)], count, acc=part).cpu().numpy())
            acc["flat"] += part["flat"]               # what the NCCL reduce does
        rows = np.concatenate(rows)
        xk, a0k, a1k = FxEngine.finish_integration(acc)
        for b in range(nb):
            assert close(rows[b], rows1[b], 2e-6), (world, b)
        assert close(xk, x1, 1e-6) and close(a0k, a01, 1e-6)      # float32 partials per segment, float64 across
    eng.close()


def test_c5_short_integrations():
    """configs[4]: bw 3.2e6, resolution 1024, 0.1 s integrations = 312 frames = 319488 samples each;
    a batch of integrations vs the oracle, and the CSV text of the rows reads back exactly."""
    S, N, nb = 319488, 1024, 24
    raw0, raw1 = synth.correlated_pair(nb * S, delay=9, seed=5)
    tau = 9 / 3.2e6
    eng = FxEngine(S, N, 4, max_blocks=nb)
    eng.set_delay(3.2e6, 1.4204e9, tau)
    x = eng.process_host(raw0, raw1, nb)
    for b in (0, 7, 23):
        sl = slice(2 * S * b, 2 * S * (b + 1))
        assert close(x[b], orc.process_block_u8(raw0[sl], raw1[sl], N, 3.2e6, 1.4204e9, tau)), b
    import io
    from effex_b200 import csvio
    back = np.loadtxt(io.BytesIO(csvio.format_rows(x)), dtype=np.complex128, delimiter=',')
    np.testing.assert_array_equal(back, x.astype(np.complex128))
    eng.close()


@pytest.mark.parametrize("S,N", [(2**16, 4096), (2**13, 1024), (2**14, 2048), (2**12, 256), (2**15, 8192)])
def test_streaming_history_equals_one_giant_block(S, N):
    """Streaming mode (PFB history carried across blocks, recording-wide mean) = the reference's
    arithmetic applied to the whole recording as ONE block; and cutting the recording into time shards
    with halos and global byte sums (what ranks do) changes nothing."""
    nb = 12
    raw0, raw1 = synth.correlated_pair(nb * S, delay=7, dc0=0.02 + 0.01j, dc1=-0.015j, seed=17)
    d0, d1 = dev(raw0), dev(raw1)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    w = orc.pfb_window(4, N)
    x0, x1 = orc.block_from_u8(raw0), orc.block_from_u8(raw1)          # ONE block: global mean
    f0, f1 = orc.spectrometer_poly(x0, 4, N, w), orc.spectrometer_poly(x1, 4, N, w)
    ref = np.fft.fftshift((f0 * np.conj(f1)).mean(axis=0))
    ref_a0 = np.fft.fftshift((abs(f0) ** 2).mean(axis=0))
    acc = eng.new_accumulators()
    eng.integrate_stream(d0, d1, acc, nb)
    x, a0, a1 = FxEngine.finish_integration(acc)
    assert acc["frames"].item() == nb * S // N
    assert close(x, ref) and close(a0, ref_a0)
    # reference mode differs (zero history + per-block mean at every block): the two semantics are distinct
    acc_ref = eng.new_accumulators()
    eng.integrate(d0, d1, acc_ref, nb)
    xr, _, _ = FxEngine.finish_integration(acc_ref)
    assert not close(xr, ref, 1e-3)
    # time shards with halos
    sums = eng.span_sums(d0, d1, nb)
    np.testing.assert_array_equal(sums, [raw0[0::2].sum(dtype=np.uint64), raw0[1::2].sum(dtype=np.uint64),
                                         raw1[0::2].sum(dtype=np.uint64), raw1[1::2].sum(dtype=np.uint64)])
    hb = 2 * 3 * N
    for world in (2, 3, 4):
        tot = eng.new_accumulators()
        for rank in range(world):
            start, count = sharding.shard_range(nb, world, rank)
            lo, hi = 2 * S * start, 2 * S * (start + count)
            h0 = d0[lo - hb:lo].contiguous() if rank else None
            h1 = d1[lo - hb:lo].contiguous() if rank else None
            part = eng.new_accumulators()
            eng.integrate_stream(d0[lo:hi], d1[lo:hi], part, count, h0, h1, sums, nb * S)
            tot["flat"] += part["flat"]
        xs, a0s, _ = FxEngine.finish_integration(tot)
        assert close(xs, ref) and close(xs, x, 2e-6) and close(a0s, a0, 2e-6), world
    eng.close()


def test_big_nbins_streaming_and_rows_walk_several_chunks_of_Z(monkeypatch):
    """8192 bins with the Z buffer limited to 64 frames: the streaming span (200 frames, halo, recording-wide
    mean) is walked in four frame chunks and the block mode in block chunks; both equal the oracle."""
    monkeypatch.setenv("EFFEX_FX_Z_ELEMS", str(64 * 8192))
    S, N, nb = 25 * 8192, 8192, 8
    raw0, raw1 = synth.correlated_pair(nb * S, delay=7, dc0=0.02 + 0.01j, dc1=-0.015j, seed=23)
    d0, d1 = dev(raw0), dev(raw1)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    w = orc.pfb_window(4, N)
    x0, x1 = orc.block_from_u8(raw0), orc.block_from_u8(raw1)          # ONE block: global mean
    f0, f1 = orc.spectrometer_poly(x0, 4, N, w), orc.spectrometer_poly(x1, 4, N, w)
    ref = np.fft.fftshift((f0 * np.conj(f1)).mean(axis=0))
    acc = eng.new_accumulators()
    eng.integrate_stream(d0, d1, acc, nb)
    x, _, _ = FxEngine.finish_integration(acc)
    assert acc["frames"].item() == nb * S // N == 200
    assert close(x, ref)
    # second half alone, with its halo and the recording-wide sums: adds up with the first half
    sums = eng.span_sums(d0, d1, nb)
    tot = eng.new_accumulators()
    half = nb // 2
    eng.integrate_stream(d0[:2 * S * half], d1[:2 * S * half], tot, half, sums=sums, total_samp=nb * S)
    hb = 2 * 3 * N
    lo = 2 * S * half
    eng.integrate_stream(d0[lo:], d1[lo:], tot, nb - half, halo0=d0[lo - hb:lo], halo1=d1[lo - hb:lo], sums=sums,
                         total_samp=nb * S)
    xt, _, _ = FxEngine.finish_integration(tot)
    assert close(xt, ref)
    # block mode: 25 frames per block -> two blocks per chunk of Z
    rows = eng.process(d0, d1, nb).cpu().numpy()
    refr = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, 0.0, 4, 0, nb)
    for b in range(nb):
        assert close(rows[b], refr[b])
    eng.close()
