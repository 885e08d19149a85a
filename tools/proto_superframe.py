#!/usr/bin/env python
"""numpy emulation of the index algebra of the two generalisations of the fused kernel
(effex_b200/csrc/fx_fused4096s.cuh, fx_bigfft.cuh), checked against numpy.fft:

  super-frames (nbins = 4096 >> LOGF): a slot of 4096 samples holds F = 2^LOGF frames; thread t owns samples
      t + 256 r, i.e. RP = 16/F positions of each frame (r = f*RP + r'); stage A = F DFTs of RP points with
      twiddle W_N^(t*k1') into tile f*RP + k1'; stages B, C = one 256-point transform per tile; tile
      f*RP + k1' then holds bins k1' + RP*(k2 + 16*k3) of frame f.
  head/tail (nbins = G*4096): head = G-point DFT over n1 of w[n1*4096 + n2], twiddle W_N^(n2*k1) -> Z[k1][n2];
      tail = 4096-point transform of every Z[k1] -> bins k1 + G*k2.

Run: python tools/proto_superframe.py   (also imported by tests/test_proto_cpu.py)"""
import numpy as np

NT = 256


def perm16(j):
    return (j >> 2) + 4 * (j & 3)


def perm_rp(rp, jj):
    if rp == 16:
        return perm16(jj)
    if rp == 8:
        return 2 * jj if jj < 4 else 2 * (jj - 4) + 1
    return jj


def row_of(logf, j):
    rp = 16 >> logf
    return (j // rp) * rp + perm_rp(rp, j % rp)


def dft_small_positions(v):
    """DFT over the last axis with the register order of the kernels: position jj holds Y[perm_rp(RP, jj)]."""
    rp = v.shape[-1]
    y = np.fft.fft(v, axis=-1)
    return np.stack([y[..., perm_rp(rp, jj)] for jj in range(rp)], axis=-1)


def tile_fft256(tile):
    """stages B and C on one tile [row = t >> 4][col = t & 15] -> X[k2 + 16*k3] (natural order)."""
    x = tile.reshape(256)                     # index t = n2*16 + n3
    return np.fft.fft(x)


def superframe_emulated(samples, logf):
    """samples: (4096,) complex = F consecutive frames of NL = 4096 >> logf -> (F, NL) spectra."""
    F, RP, NL = 1 << logf, 16 >> logf, 4096 >> logf
    t = np.arange(NT)
    v = np.stack([samples[t + 256 * r] for r in range(16)], axis=1)          # (256, 16): r = f*RP + r'
    v = v.reshape(NT, F, RP)
    v = dft_small_positions(v).reshape(NT, 16)
    tiles = np.zeros((16, 256), complex)                                       # tile index = exchange-1 row
    for j in range(16):
        row = row_of(logf, j)
        k1p = row % RP
        tiles[row, t] = v[:, j] * np.exp(-2j * np.pi * ((t * k1p) % NL) / NL)
    out = np.zeros((F, NL), complex)
    for tile in range(16):
        f, k1p = tile // RP, tile % RP
        m = np.arange(256)                                                     # m = k2 + 16*k3
        out[f, k1p + RP * m] = tile_fft256(tiles[tile])
    return out


def bigfft_emulated(w, logg):
    """w: (G*4096,) complex FIR output of one frame -> natural-order spectrum via head (Z) and tail."""
    G, N = 1 << logg, 4096
    NB = G * N
    n2 = np.arange(N)
    v = np.stack([w[g * N + n2] for g in range(G)], axis=1)                    # (4096, G) over n1
    y = dft_small_positions(v)
    z = np.zeros((G, N), complex)
    for j in range(G):
        k1 = perm_rp(G, j)
        z[k1] = y[:, j] * np.exp(-2j * np.pi * ((n2 * k1) % NB) / NB)
    out = np.zeros(NB, complex)
    for k1 in range(G):
        out[k1 + G * np.arange(N)] = np.fft.fft(z[k1])
    return out


def check(seed=0):
    rng = np.random.default_rng(seed)
    worst = 0.0
    for logf in range(5):
        x = rng.standard_normal(4096) + 1j * rng.standard_normal(4096)
        got = superframe_emulated(x, logf)
        ref = np.fft.fft(x.reshape(1 << logf, 4096 >> logf), axis=1)
        worst = max(worst, np.abs(got - ref).max() / np.abs(ref).max())
        assert sorted(row_of(logf, j) for j in range(16)) == list(range(16))
    for logg in range(1, 5):
        x = rng.standard_normal(4096 << logg) + 1j * rng.standard_normal(4096 << logg)
        got = bigfft_emulated(x, logg)
        ref = np.fft.fft(x)
        worst = max(worst, np.abs(got - ref).max() / np.abs(ref).max())
    return worst


if __name__ == "__main__":
    print("super-frame and head/tail index algebra vs numpy.fft: worst relative error %.2e" % check())
