#!/usr/bin/env python
"""numpy emulation of the fused kernel's 16x16x16 FFT dataflow (thread = axis 0).

Mirrors effex_b200/csrc/fx_fused4096s.cuh stage by stage: same register positions (digit-reversed
radix-16 outputs), same twiddle tables, same thread -> (tile, row, column) roles in the two exchanges.
The exchange ADDRESSES here are those of the first version of the kernel (flat 256-element tiles with an
XOR swizzle); the kernel now pads tile rows to 17 elements instead -- a different bijection onto shared
memory, the same dataflow.  tools/proto_superframe.py covers nbins below and above 4096.
Run: python tools/proto_fft4096.py
"""
import numpy as np

N = 4096
NT = 256


def perm16(j):
    return (j >> 2) + 4 * (j & 3)


def radix4(a0, a1, a2, a3):
    s02, d02, s13, d13 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    return s02 + s13, d02 - 1j * d13, s02 - s13, d02 + 1j * d13


def dft16_inplace(v):
    """v: (threads, 16).  Result position j holds Y[perm16(j)]."""
    v = v.copy()
    for nb in range(4):
        v[:, nb], v[:, nb + 4], v[:, nb + 8], v[:, nb + 12] = radix4(
            v[:, nb], v[:, nb + 4], v[:, nb + 8], v[:, nb + 12])
    for ka in range(4):
        for nb in range(4):
            v[:, nb + 4 * ka] *= np.exp(-2j * np.pi * nb * ka / 16)
    for ka in range(4):
        v[:, 4 * ka], v[:, 4 * ka + 1], v[:, 4 * ka + 2], v[:, 4 * ka + 3] = radix4(
            v[:, 4 * ka], v[:, 4 * ka + 1], v[:, 4 * ka + 2], v[:, 4 * ka + 3])
    return v


def tables():
    t = np.arange(NT)
    twA = np.exp(-2j * np.pi * np.outer(np.arange(16), t) / 4096)      # [k1][t]
    twB = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256)   # [k2][n3]
    return twA, twB


def fft4096_emulated(w):
    """w: (4096,) complex -> natural-order FFT, following the kernel dataflow."""
    twA, twB = tables()
    t = np.arange(NT)
    # stage A: thread t holds w[t + 256 r]
    v = np.stack([w[t + 256 * r] for r in range(16)], axis=1)
    v = dft16_inplace(v)
    X = np.zeros(4096, complex)
    for j in range(16):
        k1 = perm16(j)
        X[k1 * 256 + t] = v[:, j] * twA[k1, t]
    # exchange 1 read: thread u -> k1 = u>>4, n3 = u&15
    u = t
    v = np.stack([X[(u >> 4) * 256 + n2 * 16 + (u & 15)] for n2 in range(16)], axis=1)
    v = dft16_inplace(v)
    X2 = np.zeros(4096, complex)
    n3 = u & 15
    base = (u >> 4) * 256
    for j in range(16):
        k2 = perm16(j)
        X2[base + k2 * 16 + (n3 ^ k2)] = v[:, j] * twB[k2, n3]
    k2t = u & 15
    v = np.stack([X2[base + k2t * 16 + (m ^ k2t)] for m in range(16)], axis=1)
    v = dft16_inplace(v)
    out = np.zeros(4096, complex)
    for j in range(16):
        k3 = perm16(j)
        out[(u >> 4) + 16 * (u & 15) + 256 * k3] = v[:, j]
    return out


def check_bank_conflicts():
    """16-byte elements: each quarter-warp (8 lanes) must hit 8 distinct
    (index mod 8) groups."""
    t = np.arange(NT)
    def ok(idx):
        return all(len(set((idx[q:q + 8] % 8).tolist())) == 8 for q in range(0, NT, 8))
    res = []
    for k1 in range(16):
        res.append(ok(k1 * 256 + t))
    for n2 in range(16):
        res.append(ok((t >> 4) * 256 + n2 * 16 + (t & 15)))
    for k2 in range(16):
        res.append(ok((t >> 4) * 256 + k2 * 16 + ((t & 15) ^ k2)))
    for m in range(16):
        res.append(ok((t >> 4) * 256 + (t & 15) * 16 + (m ^ (t & 15))))
    return all(res)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.normal(size=(7, 16)) + 1j * rng.normal(size=(7, 16))
    y = dft16_inplace(x)
    ref = np.fft.fft(x, axis=1)
    assert np.allclose(y[:, [perm16(k) for k in range(16)]], ref), "dft16"
    w = rng.normal(size=N) + 1j * rng.normal(size=N)
    got = fft4096_emulated(w)
    err = np.abs(got - np.fft.fft(w)).max()
    print("fft4096 emulation max err", err)
    assert err < 1e-9
    assert check_bank_conflicts()
    print("bank-conflict-free exchanges: OK")
