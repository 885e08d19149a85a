"""effex `.csv` output (effex.py:667-693) and its reader conventions
(effex.py:785-798, post_process.py:201-219).

Line 1: `run_time:..,bandwidth:..,frequency:..,num_samp:..,resolution:..,gain:..,mode:..`
Line 2 (SPECTRUM only): fftshifted bin frequencies, `%.18e`, comma separated.
Then one line per RUN block: complex values as ` (%.18e%+.18ej)`, comma separated
(what `np.savetxt(fh, [row], delimiter=',')` writes for complex128).  float32
results are widened to float64 before formatting so the text round-trips.
"""
from __future__ import annotations

import warnings

import numpy as np


def write_metadata(path, run_time, bandwidth, frequency, num_samp, nbins, gain, mode):
    with open(path, 'w') as fh:
        fh.write(f'run_time:{run_time},bandwidth:{bandwidth},frequency:{frequency},'
                 f'num_samp:{num_samp},resolution:{nbins},gain:{gain},mode:{mode}\n')
        if mode == 'SPECTRUM':
            freqs = np.fft.fftshift(np.fft.fftfreq(nbins, d=1 / bandwidth)) + frequency
            np.savetxt(fh, [freqs], delimiter=',')
        else:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                np.savetxt(fh, [])


def format_rows_py(rows) -> str:
    """Pure-Python formatter (complex128 precision kept; used for CONTINUUM/TEST scalars)."""
    rows = np.asarray(rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    rows = rows.astype(np.complex128)
    out = []
    for r in rows:
        out.append(','.join((' (%.18e%+.18ej)' % (v.real, v.imag)) for v in r))
    return '\n'.join(out) + '\n'


_scratch = None          # per-thread reusable output buffer (no 34 MB zero-fill and no extra copies per batch)


def _format_rows_view(rows, n_threads: int = 0):
    """complex64 rows -> a uint8 numpy VIEW of the formatted text (valid until the calling thread's next
    call): the library writes into a reusable buffer and nothing is copied on the Python side."""
    import ctypes as C
    import threading
    from . import _lib
    global _scratch
    if _scratch is None:
        _scratch = threading.local()
    lib = _lib.load()
    rows = np.ascontiguousarray(rows)
    n_rows, nbins = rows.shape
    cap = lib.fx_csv_rows_bound(n_rows, nbins)
    buf = getattr(_scratch, "buf", None)
    if buf is None or buf.size < cap:
        buf = _scratch.buf = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    rc = lib.fx_csv_format_rows(rows.ctypes.data, n_rows, nbins, n_threads, buf.ctypes.data, buf.size, C.byref(n))
    if rc != 0:
        raise ValueError(f"fx_csv_format_rows failed ({rc})")
    return buf[:n.value]


def format_rows(rows, n_threads: int = 0) -> bytes:
    """complex64 rows -> the bytes np.savetxt would write, through the library's parallel C
    formatter (fx_csv_format_rows).  Other dtypes fall back to the Python formatter."""
    rows = np.asarray(rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    if rows.dtype != np.complex64:
        return format_rows_py(rows).encode()
    return _format_rows_view(rows, n_threads).tobytes()


def write_rows(fh, rows, n_threads: int = 0):
    """format + write without an intermediate bytes object"""
    rows = np.asarray(rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    if rows.dtype != np.complex64:
        fh.write(format_rows_py(rows).encode())
    else:
        fh.write(memoryview(_format_rows_view(rows, n_threads)))


def append_rows(path, rows):
    with open(path, 'ab') as fh:
        write_rows(fh, rows)


class RowWriter:
    """Appends batches of rows to `path` from a thread of its own, in the order they were put
    (the reference's `_write_data` thread, effex.py:457-460, :687-696).  A failure in the thread is
    re-raised by the next put() or by close()."""

    def __init__(self, path, depth: int = 4):
        import queue
        import threading
        self.path = path
        self.q = queue.Queue(maxsize=depth)
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            with open(self.path, 'ab') as fh:
                while True:
                    rows = self.q.get()
                    if rows is None:
                        return
                    if self.error is None:
                        write_rows(fh, rows)
        except Exception as e:          # keep draining so that put() never blocks forever
            self.error = e
            while self.q.get() is not None:
                pass

    def put(self, rows):
        if self.error is not None:
            raise self.error
        self.q.put(np.array(rows, copy=True))       # the caller may reuse its buffer

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


def read_metadata(path) -> dict:
    with open(path) as fh:
        first = fh.readline().strip()
    return dict(item.split(':') for item in first.split(','))


def read_rows(path):
    """What effex.py:798 / post_process.py:219 do."""
    meta = read_metadata(path)
    skip = 1 if meta['mode'].lower() in ('continuum', 'test') else 2
    return meta, np.loadtxt(path, dtype=np.complex128, delimiter=',', skiprows=skip, ndmin=2)
