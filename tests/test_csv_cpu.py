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
