import numpy as np

from effex_b200.ingest import RecordingReader


def test_reader_delivers_whole_blocks_in_order(tmp_path):
    S, nb = 512, 11
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=2 * S * nb + 77, dtype=np.uint8)      # ragged tail
    b = rng.integers(0, 256, size=2 * S * nb + 5, dtype=np.uint8)
    p0, p1 = tmp_path / "a.iq", tmp_path / "b.iq"
    a.tofile(p0); b.tofile(p1)
    r = RecordingReader(str(p0), str(p1), S, batch_blocks=4, skip_blocks=1)
    assert r.n_blocks == nb - 1 and len(r) == 3
    got0, got1, firsts = [], [], []
    for c0, c1, first, n in r:
        got0.append(c0.copy()); got1.append(c1.copy()); firsts.append((first, n))
    assert firsts == [(0, 4), (4, 4), (8, 2)]
    np.testing.assert_array_equal(np.concatenate(got0), a[2 * S:2 * S * nb])
    np.testing.assert_array_equal(np.concatenate(got1), b[2 * S:2 * S * nb])
    assert RecordingReader(str(p0), str(p1), S, batch_blocks=4, max_blocks=3).n_blocks == 3


def test_hostmem_bind_is_safe_without_nvml_or_gpu():
    """bind_to_gpu never widens the affinity mask and is a no-op when NVML has no answer."""
    import os
    from effex_b200 import hostmem
    if not hasattr(os, "sched_getaffinity"):
        return
    before = os.sched_getaffinity(0)
    cpus = hostmem.bind_to_gpu(0)
    after = os.sched_getaffinity(0)
    try:
        assert after <= before and set(cpus) == after
    finally:
        os.sched_setaffinity(0, before)
