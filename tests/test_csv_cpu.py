"""The C row formatter writes exactly the bytes np.savetxt writes (effex.py:693)."""
import io

import numpy as np
import pytest

from effex_b200 import csvio
from oracle import fx_oracle as orc


def _savetxt(rows):
    buf = io.StringIO()
    for r in rows:
        np.savetxt(buf, [np.asarray(r, dtype=np.complex128)], delimiter=',')
    return buf.getvalue().encode()


@pytest.mark.parametrize("shape", [(1, 1), (3, 16), (7, 1024), (33, 4096)])
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_c_formatter_matches_numpy(shape, threads):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    x = ((rng.normal(size=shape) + 1j * rng.normal(size=shape)) * 10.0 ** rng.integers(-30, 30, size=shape)
         ).astype(np.complex64)
    x.flat[0] = 0
    if x.size > 4:
        x.flat[1] = complex(-0.0, np.float32(1e-45))          # negative zero, smallest denormal
        x.flat[2] = complex(np.float32(3.4028235e38), -np.float32(1.17549435e-38))
        x.flat[3] = complex(-1.0, 1.0)
    assert csvio.format_rows(x, threads) == _savetxt(x)


def test_rows_round_trip_like_the_reference_reader(tmp_path):
    rng = np.random.default_rng(5)
    rows = (rng.normal(size=(4, 256)) + 1j * rng.normal(size=(4, 256))).astype(np.complex64)
    path = tmp_path / "v.csv"
    csvio.write_metadata(str(path), 1, 2.4e6, 1.4204e9, 4096, 256, 49.6, "SPECTRUM")
    csvio.append_rows(str(path), rows[:3])
    csvio.append_rows(str(path), rows[3])
    meta, back = csvio.read_rows(str(path))
    np.testing.assert_array_equal(back, rows.astype(np.complex128))
    assert open(path).read().startswith(orc.csv_metadata(1, 2.4e6, 1.4204e9, 4096, 256, 49.6, "SPECTRUM"))


def test_python_formatter_for_double_precision_scalars():
    v = np.array([[1.2345678901234567e-9 - 7.1e-12j]])
    assert csvio.format_rows(v) == _savetxt(v)


def test_exact_formatter_is_printf():
    """fx_csv_format_double (exact integer arithmetic, no snprintf) == '%.18e' / '%+.18e' on float32-born
    values of every exponent, on full doubles, on halfway cases (values with exactly 20 significant digits:
    round-half-even decides) and on the special values."""
    import ctypes as C
    from effex_b200 import _lib
    lib = _lib.load()
    buf = C.create_string_buffer(40)

    def fmt(v, plus):
        n = lib.fx_csv_format_double(float(v), plus, buf)
        return buf.raw[:n].decode()
    rng = np.random.default_rng(11)
    f32 = rng.integers(0, 2**32, size=60000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    f64 = rng.integers(0, 2**63, size=30000, dtype=np.uint64).view(np.float64)
    ties = [m * 2.0 ** -k for k in range(1, 70) for m in (1, 3, 5, 7, 9, 11, 1023, 4095, 99999)]
    special = [0.0, -0.0, 1.0, -1.0, 9.9999999999999999e22, 1e23, 5e-324, 1.7976931348623157e308,
               2.2250738585072014e-308, 0.1, 9.5, 99999999999999999999.0, float("inf"), float("-inf")]
    n = 0
    for arr in (f32[np.isfinite(f32)].astype(np.float64), f64[np.isfinite(f64)], ties, special):
        for i, v in enumerate(arr):
            plus = i & 1
            assert fmt(v, plus) == (("%+.18e" if plus else "%.18e") % v), repr(float(v))
            n += 1
    assert n > 80000


def test_row_writer_thread_keeps_order_and_bytes(tmp_path):
    """the writer thread (reference: effex.py:457-460, :687-696) appends batches in put() order and writes
    the same bytes as synchronous appends; the caller may reuse its buffer right after put()"""
    rng = np.random.default_rng(2)
    batches = [(rng.normal(size=(k, 64)) + 1j * rng.normal(size=(k, 64))).astype(np.complex64) for k in (3, 1, 7, 2)]
    a, b = tmp_path / "a.csv", tmp_path / "b.csv"
    for p in (a, b):
        csvio.write_metadata(str(p), 1, 2.4e6, 1.4204e9, 4096, 64, 49.6, "SPECTRUM")
    for x in batches:
        csvio.append_rows(str(a), x)
    w = csvio.RowWriter(str(b))
    scratch = np.empty((7, 64), dtype=np.complex64)
    for x in batches:
        scratch[:len(x)] = x
        w.put(scratch[:len(x)])
        scratch[:] = 0                      # reuse at once
    w.close()
    assert a.read_bytes() == b.read_bytes()
    w = csvio.RowWriter(str(tmp_path / "no_such_dir" / "c.csv"))
    with pytest.raises(OSError):
        w.put(batches[0]); w.close()
