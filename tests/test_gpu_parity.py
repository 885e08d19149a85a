"""GPU parity: the CUDA path (through the C ABI) against the float64 oracle.

Bars (BASELINE.json north_star): accumulated cross-spectra within 1e-4
relative (max-norm and L2-norm) of the float64 oracle; the
integer delay lag bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fx_oracle as orc
from effex_b200 import synth
from effex_b200.engine import FxEngine, rot_vector

pytestmark = pytest.mark.gpu

TOL = 1e-4      # north_star tolerance, float32 GPU vs float64 reference


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.complex128)
    ref = np.asarray(ref, dtype=np.complex128)
    return (np.abs(got - ref).max() / np.abs(ref).max(),
            np.linalg.norm(got - ref) / np.linalg.norm(ref))


def assert_close(got, ref, tol=TOL, what=""):
    e_max, e_l2 = rel_err(got, ref)
    assert e_max <= tol and e_l2 <= tol, f"{what}: max-norm {e_max:.3e}, L2 {e_l2:.3e} > {tol}"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_rows(raw0, raw1, S, N, bw, fc, tau, nblocks, ntaps=4):
    return orc.process_recording_u8(raw0, raw1, S, N, bw, fc, tau, ntaps, 0, nblocks)


def oracle_autos(raw0, raw1, S, N, b, ntaps=4):
    sl = slice(2 * S * b, 2 * S * (b + 1))
    return orc.auto_powers(orc.block_from_u8(raw0[sl]), orc.block_from_u8(raw1[sl]), ntaps, N,
                           orc.pfb_window(ntaps, N))


# --------------------------------------------------------------------------
# canonical config C1: S = 2^18, N = 4096, T = 4 -> fused kernel
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c1():
    S, N, nb = 2**18, 4096, 6
    raw0, raw1 = synth.correlated_pair(nb * S, delay=37, dc0=0.011 - 0.007j, dc1=-0.004 + 0.015j)
    return dict(S=S, N=N, nb=nb, raw0=raw0, raw1=raw1, bw=2.4e6, fc=1.4204e9)


@pytest.mark.parametrize("tau_samples", [0.0, 37.0])
def test_fused_canonical_vs_oracle(c1, tau_samples):
    S, N, nb = c1["S"], c1["N"], c1["nb"]
    tau = tau_samples / c1["bw"]
    eng = FxEngine(S, N, 4, max_blocks=nb)
    assert eng.fused
    eng.set_delay(c1["bw"], c1["fc"], tau)
    x, a0, a1 = eng.process(dev(c1["raw0"]), dev(c1["raw1"]), nb, autos=True)
    eng.sync()
    ref = oracle_rows(c1["raw0"], c1["raw1"], S, N, c1["bw"], c1["fc"], tau, nb)
    got = x.cpu().numpy()
    for b in range(nb):
        assert_close(got[b], ref[b], what=f"block {b}")
    r0, r1 = oracle_autos(c1["raw0"], c1["raw1"], S, N, 1)
    assert_close(a0[1].cpu().numpy(), r0, what="auto0")
    assert_close(a1[1].cpu().numpy(), r1, what="auto1")
    eng.close()


def test_fused_equals_generic_path(c1):
    S, N, nb = c1["S"], c1["N"], 3
    d0, d1 = dev(c1["raw0"]), dev(c1["raw1"])
    ef = FxEngine(S, N, 4, max_blocks=nb)
    eg = FxEngine(S, N, 4, max_blocks=nb, force_generic=True)
    assert ef.fused and not eg.fused
    xf = ef.process(d0, d1, nb).cpu().numpy()
    xg = eg.process(d0, d1, nb).cpu().numpy()
    for b in range(nb):
        assert_close(xf[b], xg[b], tol=2e-5, what=f"fused vs generic block {b}")
    ef.close(); eg.close()


def test_staggered_equals_lockstep_kernel(c1):
    """The two fused kernels (staggered/TMEM transposed FIR vs lock-step direct FIR) differ only in
    float32 summation order."""
    S, N, nb = c1["S"], c1["N"], c1["nb"]
    d0, d1 = dev(c1["raw0"]), dev(c1["raw1"])
    es = FxEngine(S, N, 4, max_blocks=nb)
    el = FxEngine(S, N, 4, max_blocks=nb, lockstep_kernel=True)
    xs, a0s, a1s = es.process(d0, d1, nb, autos=True)
    xl, a0l, a1l = el.process(d0, d1, nb, autos=True)
    for b in range(nb):
        assert_close(xs[b].cpu().numpy(), xl[b].cpu().numpy(), tol=2e-6, what=f"block {b}")
        assert_close(a0s[b].cpu().numpy(), a0l[b].cpu().numpy(), tol=2e-6, what=f"auto0 block {b}")
    es.close(); el.close()


@pytest.mark.parametrize("nb", [1, 2, 5, 149, 300])
def test_fused_segment_plans_agree(c1, nb):
    """Every block count gets its own partition of the frames into per-CTA segments; rows must not
    depend on it beyond float32 summation order, and a repeated call must be bit-identical.
    Input = 2 distinct blocks tiled."""
    S, N = c1["S"], c1["N"]
    base0, base1 = c1["raw0"][:4 * S], c1["raw1"][:4 * S]
    raw0 = np.tile(base0, (nb + 1) // 2)[:2 * S * nb]
    raw1 = np.tile(base1, (nb + 1) // 2)[:2 * S * nb]
    eng = FxEngine(S, N, 4, max_blocks=nb)
    d0, d1 = dev(raw0), dev(raw1)
    x = eng.process(d0, d1, nb).cpu().numpy()
    ref = oracle_rows(base0, base1, S, N, 2.4e6, 1.4204e9, 0.0, min(nb, 2))
    for b in range(nb):
        assert_close(x[b], ref[b % 2], what=f"nb={nb} block {b}")
        if b >= 2:      # identical input blocks, cut into segments at different frames
            assert_close(x[b], x[b - 2], tol=2e-6, what=f"nb={nb} block {b} vs {b - 2}")
    again = eng.process(d0, d1, nb).cpu().numpy()
    np.testing.assert_array_equal(x, again)         # deterministic: no atomics on the data path
    eng.close()


def test_fused_large_dc_offset():
    S, N = 2**16, 4096
    raw0, raw1 = synth.correlated_pair(S, delay=3, dc0=0.09 + 0.05j, dc1=-0.07 - 0.11j, seed=5)
    eng = FxEngine(S, N, 4)
    assert eng.fused
    x = eng.process(dev(raw0), dev(raw1), 1).cpu().numpy()[0]
    ref = orc.process_block_u8(raw0, raw1, N, 2.4e6, 1.4204e9, 0.0)
    assert_close(x, ref, what="large DC")
    # without DC removal the spike at the centre bin would dominate
    eng2 = FxEngine(S, N, 4, dc_remove=False)
    y = eng2.process(dev(raw0), dev(raw1), 1).cpu().numpy()[0]
    assert np.abs(y[N // 2]) > 20 * np.abs(x).max()
    eng.close(); eng2.close()


def test_integrate_matches_rows(c1):
    S, N, nb = c1["S"], c1["N"], c1["nb"]
    eng = FxEngine(S, N, 4, max_blocks=nb)
    d0, d1 = dev(c1["raw0"]), dev(c1["raw1"])
    rows = eng.process(d0, d1, nb).cpu().numpy().astype(np.complex128)
    acc = eng.new_accumulators()
    eng.integrate(d0[:2 * S * 4], d1[:2 * S * 4], acc, 4)
    eng.integrate(d0[2 * S * 4:], d1[2 * S * 4:], acc, nb - 4)     # accumulates across calls
    eng.sync()
    assert acc["frames"].item() == nb * (S // N)
    x, a0, a1 = FxEngine.finish_integration(acc)
    assert_close(x, rows.mean(axis=0), tol=1e-6, what="integrate vs mean of rows")
    ref = oracle_rows(c1["raw0"], c1["raw1"], S, N, c1["bw"], c1["fc"], 0.0, nb).mean(axis=0)
    assert_close(x, ref, what="integrate vs oracle")
    eng.close()


def test_process_host_pipeline(c1):
    S, N, nb = c1["S"], c1["N"], c1["nb"]
    eng = FxEngine(S, N, 4, max_blocks=4)          # forces several chunks (6 blocks, <= 4 per chunk)
    x = eng.process_host(c1["raw0"], c1["raw1"], nb)
    ref = eng.process(dev(c1["raw0"])[:2 * S * 4], dev(c1["raw1"])[:2 * S * 4], 4).cpu().numpy()
    np.testing.assert_array_equal(x[:4], ref)
    full = oracle_rows(c1["raw0"], c1["raw1"], S, N, c1["bw"], c1["fc"], 0.0, nb)
    for b in range(nb):
        assert_close(x[b], full[b], what=f"host pipeline block {b}")
    eng.close()


# --------------------------------------------------------------------------
# other shapes (generic kernels): C5-like N=1024, ragged S, small N, T up to 32
# --------------------------------------------------------------------------
@pytest.mark.parametrize("S,N,T", [(319488, 1024, 4), (4099, 2048, 2), (2**14, 256, 4), (2**15, 2048, 4),
                                   (2**16, 512, 8), (3000, 64, 32), (2**17, 8192, 4)])
def test_generic_shapes_vs_oracle(S, N, T):
    nb = 2
    raw0, raw1 = synth.correlated_pair(nb * S, delay=2, dc0=0.02j, dc1=-0.01, seed=11)
    w = orc.pfb_window(T, N)
    eng = FxEngine(S, N, T, max_blocks=nb, window=w)
    tau = 2 / 3.2e6
    eng.set_delay(3.2e6, 1.0e8, tau)
    x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    ref = orc.process_recording_u8(raw0, raw1, S, N, 3.2e6, 1.0e8, tau, T, 0, nb)
    for b in range(nb):
        assert_close(x[b], ref[b], what=f"S={S} N={N} T={T} block {b}")
    eng.close()


# --------------------------------------------------------------------------
# nbins < 4096 on the fused kernel: F = 4096/N frames per 4096-sample super-frame.
# Shapes cover whole super-frames, a partial last super-frame (P mod F != 0), a block shorter than
# one super-frame, many short blocks (segments that end inside a CTA's run) and S not a multiple of N.
# --------------------------------------------------------------------------
@pytest.mark.parametrize("S,N,nb", [(2**16, 2048, 3), (2**16, 1024, 3), (2**15, 512, 3), (2**14, 256, 3),
                                    (319488, 1024, 2), (7 * 1024, 1024, 5), (3 * 1024, 1024, 4),
                                    (11 * 256 + 8, 256, 6), (5 * 2048, 2048, 9), (2**13, 1024, 700)])
def test_fused_small_nbins_vs_oracle_and_generic(S, N, nb):
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.013 - 0.02j, dc1=-0.006 + 0.004j, seed=5)
    bw, fc, tau = 2.4e6, 1.4204e9, 3 / 2.4e6
    eng = FxEngine(S, N, 4, max_blocks=nb)
    assert eng.fused
    eng.set_delay(bw, fc, tau)
    x, a0, a1 = eng.process(dev(raw0), dev(raw1), nb, autos=True)
    x, a0, a1 = x.cpu().numpy(), a0.cpu().numpy(), a1.cpu().numpy()
    gen = FxEngine(S, N, 4, max_blocks=nb, force_generic=True)
    assert not gen.fused
    gen.set_delay(bw, fc, tau)
    xg, g0, g1 = gen.process(dev(raw0), dev(raw1), nb, autos=True)
    xg, g0, g1 = xg.cpu().numpy(), g0.cpu().numpy(), g1.cpu().numpy()
    check = range(nb) if nb <= 9 else [0, 1, nb // 2, nb - 2, nb - 1]
    ref = {b: orc.process_recording_u8(raw0, raw1, S, N, bw, fc, tau, 4, b, 1)[0] for b in check}
    for b in check:
        assert_close(x[b], ref[b], what=f"S={S} N={N} block {b} fused vs oracle")
        r0, r1 = oracle_autos(raw0, raw1, S, N, b)
        assert np.abs(a0[b] - r0).max() <= TOL * r0.max() and np.abs(a1[b] - r1).max() <= TOL * r1.max()
    # every block against the generic kernels (independent code, same arithmetic type)
    assert np.abs(x - xg).max() <= 2e-6 * np.abs(xg).max()
    assert np.abs(a0 - g0).max() <= 2e-6 * g0.max() and np.abs(a1 - g1).max() <= 2e-6 * g1.max()
    eng.close(); gen.close()


@pytest.mark.parametrize("S,N", [(6 * 1024, 1024), (2**15, 2048), (9 * 256, 256)])
def test_fused_small_nbins_integrate_and_host_pipeline(S, N):
    """fx_integrate and fx_process_host on the fused kernel below 4096 bins (P mod F != 0 for two shapes)."""
    nb = 7
    raw0, raw1 = synth.correlated_pair(nb * S, delay=2, dc0=0.01 - 0.01j, dc1=0.02j, seed=13)
    eng = FxEngine(S, N, 4, max_blocks=3)
    assert eng.fused
    ref = oracle_rows(raw0, raw1, S, N, 2.4e6, 1.4204e9, 0.0, nb)
    x = eng.process_host(raw0, raw1, nb)                      # 3 chunks of <= 3 blocks
    for b in range(nb):
        assert_close(x[b], ref[b], what=f"host pipeline N={N} block {b}")
    acc = eng.new_accumulators()
    d0, d1 = dev(raw0), dev(raw1)
    for b0 in range(0, nb, 3):
        n = min(3, nb - b0)
        eng.integrate(d0[2 * S * b0:2 * S * (b0 + n)], d1[2 * S * b0:2 * S * (b0 + n)], acc, n)
    eng.sync()
    assert acc["frames"].item() == nb * (S // N)
    xi, _, _ = FxEngine.finish_integration(acc)
    assert_close(xi, ref.mean(axis=0), what=f"integrate N={N} vs oracle")
    eng.close()


# --------------------------------------------------------------------------
# nbins = G * 4096 (G = 2..16): head kernel -> Z -> tail kernel (fx_bigfft.cuh)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("S,N,nb", [(3 * 8192, 8192, 3), (5 * 16384 + 24, 16384, 2), (2 * 32768, 32768, 2),
                                    (4 * 65536, 65536, 2), (65536, 65536, 5), (2**18, 8192, 40)])
def test_big_nbins_vs_oracle_and_generic(S, N, nb):
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.013 - 0.02j, dc1=-0.006 + 0.004j, seed=6)
    bw, fc, tau = 2.4e6, 1.4204e9, 3 / 2.4e6
    eng = FxEngine(S, N, 4, max_blocks=nb)
    eng.set_delay(bw, fc, tau)
    eng.reset_counters()
    x, a0, a1 = eng.process(dev(raw0), dev(raw1), nb, autos=True)
    assert eng.kernel_launches() <= 8, "expected the head/tail kernels, not one launch per FFT pass"
    x, a0, a1 = x.cpu().numpy(), a0.cpu().numpy(), a1.cpu().numpy()
    gen = FxEngine(S, N, 4, max_blocks=nb, force_generic=True)
    gen.set_delay(bw, fc, tau)
    xg, g0, g1 = gen.process(dev(raw0), dev(raw1), nb, autos=True)
    xg, g0, g1 = xg.cpu().numpy(), g0.cpu().numpy(), g1.cpu().numpy()
    check = range(nb) if nb <= 5 else [0, 1, nb // 2, nb - 1]
    for b in check:
        ref = orc.process_recording_u8(raw0, raw1, S, N, bw, fc, tau, 4, b, 1)[0]
        assert_close(x[b], ref, what=f"S={S} N={N} block {b} vs oracle")
        r0, r1 = oracle_autos(raw0, raw1, S, N, b)
        assert np.abs(a0[b] - r0).max() <= TOL * r0.max() and np.abs(a1[b] - r1).max() <= TOL * r1.max()
    # against the unfused kernels (independent code, same arithmetic type, different summation orders)
    assert np.abs(x - xg).max() <= 1e-5 * np.abs(xg).max()
    assert np.abs(a0 - g0).max() <= 1e-5 * g0.max() and np.abs(a1 - g1).max() <= 1e-5 * g1.max()
    # rows + float64 accumulators in one call, and accumulators alone (fx_integrate)
    acc = eng.new_accumulators()
    xa = eng.process(dev(raw0), dev(raw1), nb, acc=acc).cpu().numpy()
    np.testing.assert_array_equal(xa, x)
    acc2 = eng.new_accumulators()
    eng.integrate(dev(raw0), dev(raw1), acc2, nb)
    eng.sync()
    assert acc["frames"].item() == acc2["frames"].item() == nb * (S // N)
    for a in (acc, acc2):
        xi, i0, i1 = FxEngine.finish_integration(a, rot=rot_vector(N, bw, fc, tau))
        assert_close(xi, x.astype(np.complex128).mean(axis=0), tol=2e-6, what="integrated cross-spectrum vs mean of rows")
        assert_close(i0, a0.astype(np.float64).mean(axis=0), tol=2e-6, what="integrated auto0")
    eng.close(); gen.close()


def test_big_nbins_several_chunks_of_Z():
    """More blocks than one Z buffer (1 GiB) holds: 20 blocks of 2^22 samples at 65536 bins -> two chunks."""
    S, N, nb = 2**22, 65536, 20
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=3)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    for b in (0, 17):
        ref = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, 0.0, 4, b, 1)[0]
        assert_close(x[b], ref, what=f"block {b} vs oracle")
    for b in range(3, nb):                       # tiled input: block b repeats block b mod 3
        assert_close(x[b], x[b % 3], tol=2e-6, what=f"block {b} vs block {b % 3}")
    eng.close()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_golden_reference_fixture(golden_dir, tag):
    """Fixtures made by running the reference's own effex.py (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, f"ref_case_{tag}.npz"))
    S, N, bw, fc = int(g["S"]), int(g["N"]), float(g["bw"]), float(g["fc"])
    eng = FxEngine(S, N, 4)
    d0, d1 = dev(g["raw0"]), dev(g["raw1"])
    assert_close(eng.process(d0, d1, 1).cpu().numpy()[0], g["xspec_tau0"], what="xspec tau=0")
    eng.set_delay(bw, fc, float(g["calibrated_delay"]))
    assert_close(eng.process(d0, d1, 1).cpu().numpy()[0], g["xspec_cal"], what="xspec calibrated")
    assert_close(eng.pfb(d0).cpu().numpy(), g["spec0"], what="spectrometer (u8)")
    x0 = orc.block_from_u8(g["raw0"])
    assert_close(eng.pfb(x0).cpu().numpy(), g["spec0"], what="spectrometer (c64)")
    eng.close()


# --------------------------------------------------------------------------
# delay calibration: integer lag exact
# --------------------------------------------------------------------------
@pytest.mark.parametrize("nblk", [1, 4])
def test_lag_u8_exact_c2(nblk):
    S = 2**18
    raw0, raw1 = synth.correlated_pair(nblk * S, delay=37, seed=123)
    eng = FxEngine(S, 4096, 4, max_blocks=nblk)
    n, imax, p, q, r = eng.lag(dev(raw0), dev(raw1), nblk)
    blocks0 = [orc.block_from_u8(raw0[2 * S * b:2 * S * (b + 1)]) for b in range(nblk)]
    blocks1 = [orc.block_from_u8(raw1[2 * S * b:2 * S * (b + 1)]) for b in range(nblk)]
    rn, rimax, rp, rq, rr = orc.accumulated_lag_search(blocks0, blocks1)
    assert n - imax == 37 and imax == rimax            # bit-exact integer lag
    # neighbours: float32 transforms of 2^19 points carry ~1e-6 of the PEAK as absolute error
    np.testing.assert_allclose([p, q, r], [rp, rq, rr], rtol=1e-5, atol=3e-6 * rq)
    eng.close()


@pytest.mark.parametrize("n", [3 + 2**12, 2**18, 1000])
@pytest.mark.parametrize("offset", [-2000, -1001, -1, 0, 1, 999, 2000])
def test_lag_c64_exact(n, offset):
    if abs(offset) >= n // 2:
        pytest.skip("a circular roll by more than n/2 is the shorter roll the other way")
    iq0, iq1 = synth.rolled_pair(n, offset)
    eng = FxEngine(n, 8, 1)
    gn, imax, p, q, r = eng.lag(iq0, iq1)
    rn, rimax, rp, rq, rr = orc.lag_search(iq0, iq1)
    assert imax == rimax and gn - imax == offset
    np.testing.assert_allclose([p, q, r], [rp, rq, rr], rtol=1e-5, atol=3e-6 * rq)
    eng.close()


@pytest.mark.parametrize("n,nblk", [(3 + 2**12, 3), (2**16, 5), (2**18, 2), (2**13 + 8, 1)])
def test_lag_head_tail_kernels_equal_the_generic_passes(n, nblk):
    """The lag search on the fused kernel's FFT machinery (fx_lag.cuh: head -> Z -> tail, blocks accumulated in
    registers) against the unfused Stockham passes, from raw bytes and from complex input; and the two
    halves of the search (accumulate, finish) against the one call."""
    raw0, raw1 = synth.correlated_pair(nblk * n, delay=-11, seed=5, dc0=0.03, dc1=0.01j)
    d0, d1 = dev(raw0), dev(raw1)
    fast, slow = FxEngine(n, 8, 1, max_blocks=nblk), FxEngine(n, 8, 1, max_blocks=nblk, force_generic=True)
    a, b = fast.lag(d0, d1, nblk), slow.lag(d0, d1, nblk)
    assert a[1] == b[1] and a[0] - a[1] == -11
    np.testing.assert_allclose(a[2:], b[2:], rtol=1e-5, atol=3e-6 * b[3])
    xa, xb = fast.lag_accumulate(d0, d1, nblk), slow.lag_accumulate(d0, d1, nblk)
    scale = float(xb.abs().max())
    assert float((xa - xb).abs().max()) <= 2e-5 * scale
    half = nblk // 2
    if half:
        x2 = fast.lag_accumulate(d0[:2 * n * half], d1[:2 * n * half], half)
        fast.lag_accumulate(d0[2 * n * half:], d1[2 * n * half:], nblk - half, xacc=x2, first=False)
        assert float((x2 - xa).abs().max()) <= 2e-6 * scale
    assert fast.lag_finish(xa)[1] == a[1]
    x0 = torch.from_numpy(orc.block_from_u8(raw0[:2 * n]).astype(np.complex64)).cuda()
    x1 = torch.from_numpy(orc.block_from_u8(raw1[:2 * n]).astype(np.complex64)).cuda()
    c, d = fast.lag(x0, x1), slow.lag(x0, x1)
    assert c[1] == d[1]
    np.testing.assert_allclose(c[2:], d[2:], rtol=1e-5, atol=3e-6 * d[3])
    fast.close(); slow.close()


def test_errors_are_value_errors():
    with pytest.raises(ValueError):
        FxEngine(2**18, 4096, 33)
    with pytest.raises(ValueError):
        FxEngine(2**18, 70000, 4)              # nbins outside [8, 65536]
    eng = FxEngine(2**14, 1024, 4, max_blocks=1)
    raw = torch.zeros(4 * 2**14, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        eng.process(raw, raw, 2)               # exceeds max_blocks
    with pytest.raises(ValueError):
        eng.process(raw.cpu(), raw.cpu(), 1)   # host tensor on the device entry point
    eng.close()


def test_unaligned_inputs_take_the_generic_kernels(c1):
    """The fused kernel needs 16-byte aligned blocks (TMA bulk copies); a misaligned view must be
    detected and routed to the generic kernels, with the same answer."""
    S, N = c1["S"], c1["N"]
    buf0 = torch.zeros(2 * S + 64, dtype=torch.uint8, device="cuda")
    buf1 = torch.zeros(2 * S + 64, dtype=torch.uint8, device="cuda")
    v0, v1 = buf0[2:2 + 2 * S], buf1[6:6 + 2 * S]
    v0.copy_(torch.from_numpy(c1["raw0"][:2 * S])); v1.copy_(torch.from_numpy(c1["raw1"][:2 * S]))
    assert v0.data_ptr() % 16 != 0
    eng = FxEngine(S, N, 4)
    before = eng.kernel_launches()
    x = eng.process(v0.contiguous() if not v0.is_contiguous() else v0, v1, 1).cpu().numpy()[0]
    ref = orc.process_block_u8(c1["raw0"][:2 * S], c1["raw1"][:2 * S], N, 2.4e6, 1.4204e9, 0.0)
    assert_close(x, ref, what="unaligned input")
    assert eng.kernel_launches() - before > 4          # FIR x2, FFT x2, X-engine, ... not the 3-launch fused path
    eng.close()


def test_many_small_blocks_and_argument_errors():
    """More blocks than one grid dimension holds (65535), tiny shapes; zero blocks is an error."""
    S, N, T, nb = 256, 64, 4, 70000
    rng = np.random.default_rng(3)
    base0 = rng.integers(0, 256, size=2 * S * 7, dtype=np.uint8)
    base1 = rng.integers(0, 256, size=2 * S * 7, dtype=np.uint8)
    raw0 = np.tile(base0, nb // 7 + 1)[:2 * S * nb]
    raw1 = np.tile(base1, nb // 7 + 1)[:2 * S * nb]
    eng = FxEngine(S, N, T, max_blocks=nb)
    x = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    ref = orc.process_recording_u8(base0, base1, S, N, 2.4e6, 1.4204e9, 0.0, T, 0, 7)
    for b in (0, 6, 7, 65534, 65535, 65536, nb - 1):
        assert_close(x[b], ref[b % 7], what=f"block {b}")
    with pytest.raises(ValueError):
        eng.process(dev(raw0), dev(raw1), 0)
    eng.close()


def test_no_dc_removal_generic_and_fused_agree_with_oracle():
    """dc_remove=0: x = b/127.5 - 1 without the mean subtraction (the unpack of pyrtlsdr alone)."""
    for S, N in ((2**16, 4096), (2**14, 512)):
        raw0, raw1 = synth.correlated_pair(S, delay=1, dc0=0.03, dc1=0.02j, seed=8)
        eng = FxEngine(S, N, 4, dc_remove=False)
        x = eng.process(dev(raw0), dev(raw1), 1).cpu().numpy()[0]
        w = orc.pfb_window(4, N)
        ref = orc.pfb_xcorr(orc.unpack_iq(raw0), orc.unpack_iq(raw1), 4, N, w, 2.4e6, 1.4204e9, 0.0)
        assert_close(x, ref, what=f"no DC removal N={N}")
        eng.close()


# --------------------------------------------------------------------------
# the previous-generation kernels kept behind environment switches (read once per process, hence a subprocess):
# the two-phase head_kernel in reference mode (EFFEX_FX_HEAD2=0) and the shared-memory Stockham lag head
# (EFFEX_FX_LAG_HEAD2=0) must keep giving the persistent / register kernels' results
# --------------------------------------------------------------------------
_FALLBACK_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from effex_b200 import synth
from effex_b200.engine import FxEngine
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
out = {}
for N, P, nb in ((8192, 5, 3), (65536, 4, 2)):
    S = P * N
    raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.013 - 0.02j, seed=N)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    out["x%%d" %% N] = eng.process(dev(raw0), dev(raw1), nb).cpu().numpy()
    eng.close()
n, nblk = 2**16, 2
raw0, raw1 = synth.correlated_pair(nblk * n, delay=-7, seed=9)
eng = FxEngine(n, 8, 1, max_blocks=nblk)
r = eng.lag(dev(raw0), dev(raw1), nblk)
out["lag"] = np.array([r[0], r[1]] + list(r[2:5]), dtype=np.float64)
eng.close()
np.savez(sys.argv[1], **out)
"""


def test_fallback_kernels_behind_env_switches(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("new", {}), ("old", {"EFFEX_FX_HEAD2": "0", "EFFEX_FX_LAG_HEAD2": "0"})):
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", _FALLBACK_CHILD % root, path], env=dict(os.environ, **env), check=True,
                       timeout=600)
        res[tag] = np.load(path)
    for k in ("x8192", "x65536"):
        assert np.abs(res["new"][k] - res["old"][k]).max() <= 2e-6 * np.abs(res["old"][k]).max(), k
    assert res["new"]["lag"][0] - res["new"]["lag"][1] == -7
    np.testing.assert_array_equal(res["new"]["lag"][:2], res["old"]["lag"][:2])
    np.testing.assert_allclose(res["new"]["lag"][2:], res["old"]["lag"][2:], rtol=1e-5)


def test_big_nbins_unaligned_input_and_ragged_num_samp():
    """The persistent head kernel reads through 16-byte TMA bulk copies; inputs that are not 16-byte aligned, or a
    num_samp that is not a multiple of 8, must take the two-phase kernel and give the same rows."""
    N, nb = 8192, 3
    for S, shift in ((4 * N, 2), (4 * N + 6, 0)):
        raw0, raw1 = synth.correlated_pair(nb * S, delay=3, dc0=0.013 - 0.02j, seed=21)
        pad0, pad1 = torch.zeros(2 * nb * S + 64, dtype=torch.uint8, device="cuda"), torch.zeros(2 * nb * S + 64, dtype=torch.uint8, device="cuda")
        d0, d1 = pad0[shift:shift + 2 * nb * S], pad1[shift:shift + 2 * nb * S]
        d0.copy_(dev(raw0)); d1.copy_(dev(raw1))
        assert shift == 0 or d0.data_ptr() % 16 != 0
        eng = FxEngine(S, N, 4, max_blocks=nb)
        x = eng.process(d0, d1, nb).cpu().numpy()
        ref = orc.process_recording_u8(raw0, raw1, S, N, 2.4e6, 1.4204e9, 0.0, 4, 0, nb)
        for b in range(nb):
            assert_close(x[b], ref[b], what=f"S={S} shift={shift} block {b} vs oracle")
        eng.close()


def test_lag_graph_replay_follows_the_data_and_survives_plan_eviction():
    """fx_lag_u8 replays a CUDA graph when it is called again with the same buffers: the replay must read the buffers'
    CURRENT contents, and must be dropped when other calls have recycled the segment plans its launches point at."""
    S, N, nb = 2**16, 4096, 6
    eng = FxEngine(S, N, 4, max_blocks=nb)
    raws = {d: synth.correlated_pair(nb * S, delay=d, seed=40 + d) for d in (11, -23)}
    d0, d1 = dev(raws[11][0]).clone(), dev(raws[11][1]).clone()
    for _ in range(4):                                   # launch by launch, capture, replay, replay
        r = eng.lag(d0, d1, 2)
        assert r[0] - r[1] == 11
    first = r
    d0.copy_(dev(raws[-23][0])); d1.copy_(dev(raws[-23][1]))      # same buffers, new contents
    r = eng.lag(d0, d1, 2)
    assert r[0] - r[1] == -23
    for k in range(1, nb + 1):                           # six other shapes through the 4-slot plan cache
        eng.process(d0, d1, k)
    r = eng.lag(d0, d1, 2)
    assert r[0] - r[1] == -23
    d0.copy_(dev(raws[11][0])); d1.copy_(dev(raws[11][1]))
    for _ in range(3):
        r = eng.lag(d0, d1, 2)
        assert r[:2] == first[:2]
        np.testing.assert_allclose(r[2:], first[2:], rtol=1e-6)
    # a different block count on the same buffers is a different graph
    r1 = eng.lag(d0, d1, 1)
    assert r1[0] - r1[1] == 11
    eng.close()
