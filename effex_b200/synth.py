"""Synthetic RTL-SDR style inputs (uint8 interleaved I,Q) for the BASELINE configs.

Generators follow SURVEY.md §8(d); the base seed 77777 is the one the
reference's tests use (`tests/test_effex.py:10`).  Host-side numpy only: these
make INPUTS, they are not part of the measured path.
"""
from __future__ import annotations

import numpy as np

SEED = 77777


def quantize_u8(x: np.ndarray) -> np.ndarray:
    """complex -> uint8[2n] interleaved I,Q, inverse of pyrtlsdr's
    `packed_bytes_to_iq` (b/127.5 - 1) with rounding and clipping."""
    out = np.empty(2 * len(x), dtype=np.uint8)
    out[0::2] = np.clip(np.rint(127.5 + 127.5 * x.real), 0, 255)
    out[1::2] = np.clip(np.rint(127.5 + 127.5 * x.imag), 0, 255)
    return out


def _cnoise(rng, n, sigma):
    return rng.normal(size=n, scale=sigma) + 1j * rng.normal(size=n, scale=sigma)


def correlated_pair(n_samples: int, delay: int = 37, sigma_g: float = 0.25, sigma_n: float = 0.1,
                    seed: int = SEED, dc0: complex = 0.0, dc1: complex = 0.0):
    """C1/C2: common signal g ~ CN(0, sigma_g/component) seen by both channels,
    ch1 lagging ch0 by `delay` samples, plus independent receiver noise.
    Returns (raw0, raw1) uint8[2*n_samples]."""
    rng = np.random.default_rng(seed)
    pad = abs(delay)
    g = _cnoise(rng, n_samples + pad, sigma_g)
    if delay >= 0:
        s0, s1 = g[delay:delay + n_samples], g[:n_samples]
    else:
        s0, s1 = g[:n_samples], g[pad:pad + n_samples]
    x0 = s0 + _cnoise(rng, n_samples, sigma_n) + dc0
    x1 = s1 + _cnoise(rng, n_samples, sigma_n) + dc1
    return quantize_u8(x0), quantize_u8(x1)


def rolled_pair(n_samples: int, offset: int, scale: float = 0.1, seed: int = SEED):
    """The reference's delay-test input (`tests/test_effex.py:95-101`): complex
    Gaussian noise and a circularly rolled copy, as complex128 (un-quantised)."""
    rng = np.random.default_rng(seed)
    iq0 = _cnoise(rng, n_samples, scale)
    return iq0, np.roll(iq0, offset)


def complex_sinusoid(num_samp: int, rate: float, freq: float) -> np.ndarray:
    """`gen_complex_sinusoid` of tests/test_effex.py:31-41 (noiseless)."""
    t = np.linspace(0, num_samp / rate, num=num_samp)
    return np.cos(2.0 * np.pi * freq * t) + 1j * np.sin(2.0 * np.pi * freq * t)


def hi_line_pair(n_samples: int, bandwidth: float = 2.4e6, f0: float = 5752.0,
                 sigma_f: float = 20e3, line_rms: float = 0.03, sigma_n: float = 0.1,
                 tone_amp: float = 0.02, seed: int = SEED + 3):
    """C3: independent noise per channel + a COMMON Gaussian-profile line at
    baseband f0 (made by shaping white noise in the frequency domain over the
    whole block) + an optional pure tone as a sharp bin check."""
    rng = np.random.default_rng(seed)
    white = _cnoise(rng, n_samples, 1.0)
    f = np.fft.fftfreq(n_samples, d=1.0 / bandwidth)
    shape = np.exp(-0.5 * ((f - f0) / sigma_f) ** 2)
    line = np.fft.ifft(np.fft.fft(white) * shape)
    line *= line_rms / np.sqrt(np.mean(np.abs(line) ** 2))
    if tone_amp:
        t = np.arange(n_samples) / bandwidth
        line = line + tone_amp * np.exp(2j * np.pi * f0 * t)
    x0 = line + _cnoise(rng, n_samples, sigma_n)
    x1 = line + _cnoise(rng, n_samples, sigma_n)
    return quantize_u8(x0), quantize_u8(x1)


def tiled_recording(n_blocks: int, num_samp: int, base_blocks: int = 8, delay: int = 37,
                    seed: int = SEED):
    """A long recording made by tiling `base_blocks` freshly generated blocks
    (C4 is 17 GB/channel; tiling keeps generation time bounded while every
    block still goes through the full path)."""
    base0, base1 = correlated_pair(base_blocks * num_samp, delay=delay, seed=seed)
    reps = -(-n_blocks // base_blocks)
    raw0 = np.tile(base0, reps)[: 2 * num_samp * n_blocks]
    raw1 = np.tile(base1, reps)[: 2 * num_samp * n_blocks]
    return raw0, raw1


class TiledRecording:
    """A long recording that is `base_blocks` fresh blocks repeated (like tiled_recording) WITHOUT being
    materialised: contiguous slices are views into a window of the periodic pattern, so a one-hour
    recording (17 GB per channel) costs `window_blocks + base_blocks` blocks of (pinned, when CUDA is there)
    host memory.  Supports what Correlator.run_recording needs: `.size`, and `rec[lo:hi]` with hi - lo up to
    window_blocks blocks."""

    def __init__(self, base: np.ndarray, n_blocks: int, num_samp: int, window_blocks: int = 64):
        self.block_bytes = 2 * int(num_samp)
        self.period = base.size
        if self.period % self.block_bytes:
            raise ValueError("base must hold whole blocks")
        self.size = self.block_bytes * int(n_blocks)
        self.dtype = base.dtype
        reps = -(-(window_blocks * self.block_bytes + self.period) // self.period) + 1
        buf = np.tile(base, reps)
        try:
            import torch
            if torch.cuda.is_available():
                pinned = torch.empty(buf.size, dtype=torch.uint8).pin_memory().numpy()
                pinned[:] = buf
                buf = pinned
        except Exception:
            pass
        self._buf = buf

    def __len__(self):
        return self.size

    def __getitem__(self, key):
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise TypeError("TiledRecording supports contiguous slices only")
        lo, hi, _ = key.indices(self.size)
        off = lo % self.period
        if off + (hi - lo) > self._buf.size:
            raise ValueError("slice longer than the window this recording was built with")
        return self._buf[off:off + max(hi - lo, 0)]


def tiled_recording_lazy(n_blocks: int, num_samp: int, base_blocks: int = 8, delay: int = 37, seed: int = SEED,
                         window_blocks: int = 64):
    """tiled_recording as two TiledRecording views (same bytes, no multi-GB copies)."""
    base0, base1 = correlated_pair(base_blocks * num_samp, delay=delay, seed=seed)
    return (TiledRecording(base0, n_blocks, num_samp, window_blocks),
            TiledRecording(base1, n_blocks, num_samp, window_blocks))
