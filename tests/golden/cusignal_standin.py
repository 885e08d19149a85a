"""An independent stand-in for `cusignal.filtering.channelize_poly`, written from the STRUCTURE of cuSignal's
CUDA kernel (`_cupy_channelizer_MxM` in cusignal/filtering/_channelizer.cu + the Python wrapper in
cusignal/filtering/channelize_poly.py), thread by thread, with plain loops.  It shares no code and no
vectorised shortcut with `oracle/fx_oracle.py::channelize_poly`, so it pins that function's tap order,
sample order, conjugations, zero fill and output transpose from outside:

  wrapper   n_taps = int(len(h)/n_chans) (> 32 raises), n_pts = int(len(x)/n_chans), y = empty((n_pts, n_chans)),
            kernel, then `conj(fft(y)).T`
  kernel    a tile of M x M threads (M = 8, 16 or 32, the smallest that holds n_taps); thread (tx, ty):
              s_h[tx][ty]  = conj(h[ty*n_chans + btx])            btx = blockIdx.x*M + tx   (0 when out of range)
              for bid = blockIdx.y; bid < n_pts; bid += gridDim.y:
                  if bid >= n_taps:  s_reg[tx][n_taps-1-ty] = conj(x[(bid-n_taps+1+ty)*n_chans + (n_chans-1-btx)])
                  else:              s_reg[tx][bid-ty]      = conj(x[ty*n_chans + (n_chans-1-btx)])  (ty <= bid),
                                     every other slot 0
                  y[bid*n_chans + blockIdx.x*M + ty] = sum over tx of s_h[ty][tx] * s_reg[ty][tx]   (tile reduce)

Used as the shim in make_golden.py (so the fixtures come from the reference's own code around a routine that
is NOT the oracle's) and by tests/test_oracle_pins.py.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np


def channelize_poly(x, h, n_chans):
    x = np.asarray(x)
    h = np.asarray(h)
    dtype = np.promote_types(np.promote_types(x.dtype, h.dtype), np.complex64)
    x = x.astype(dtype)
    h = h.astype(dtype)
    n_chans = int(n_chans)
    n_taps = int(len(h) / n_chans)
    if n_taps > 32:
        raise NotImplementedError(
            "The number of calculated taps ({}) in each filter is currently capped at 32".format(n_taps))
    n_pts = int(len(x) / n_chans)
    M = 8 if n_taps <= 8 else 16 if n_taps <= 16 else 32
    y = np.empty(n_pts * n_chans, dtype=dtype)
    grid_x = -(-n_chans // M)
    for block_x in range(grid_x):
        s_h = [[0j] * M for _ in range(M)]
        for tx in range(M):
            for ty in range(M):
                btx = block_x * M + tx
                if btx < n_chans and ty < n_taps:
                    s_h[tx][ty] = np.conj(h[ty * n_chans + btx])
        for bid in range(n_pts):
            s_reg = [[0j] * M for _ in range(M)]
            for tx in range(M):
                btx = block_x * M + tx
                for ty in range(M):
                    if bid >= n_taps:
                        if btx < n_chans and ty < n_taps:
                            s_reg[tx][(n_taps - 1) - ty] = np.conj(
                                x[((bid - n_taps + 1) + ty) * n_chans + (n_chans - 1 - btx)])
                    else:
                        if btx < n_chans and ty <= bid:
                            s_reg[tx][bid - ty] = np.conj(x[ty * n_chans + (n_chans - 1 - btx)])
            for ty in range(M):                      # the thread row that owns channel block_x*M + ty
                chan = block_x * M + ty
                if chan >= n_chans:
                    continue
                vv = 0j
                for tx in range(M):                  # tile reduce over the tap / history slot
                    vv += s_h[ty][tx] * s_reg[ty][tx]
                y[bid * n_chans + chan] = vv
    y = y.reshape(n_pts, n_chans)
    return np.conj(np.fft.fft(y)).T
