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


def test_stream_reader_over_fifos(tmp_path):
    """two FIFOs fed in irregular pieces by writer threads (what two `rtl_sdr -` processes look like): whole
    blocks only, both channels in step, in order; the ragged tail of the shorter stream is dropped"""
    import os
    import threading
    from effex_b200.ingest import StreamReader, open_reader
    S, nb = 256, 9
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=2 * S * nb + 100, dtype=np.uint8)
    b = rng.integers(0, 256, size=2 * S * (nb + 2), dtype=np.uint8)
    paths = [str(tmp_path / "c0.fifo"), str(tmp_path / "c1.fifo")]
    for p in paths:
        os.mkfifo(p)

    def feed(path, data, piece):
        with open(path, 'wb', buffering=0) as f:
            for i in range(0, len(data), piece):
                f.write(data[i:i + piece].tobytes())
    ths = [threading.Thread(target=feed, args=(paths[0], a, 777)), threading.Thread(target=feed, args=(paths[1], b, 1300))]
    for t in ths:
        t.start()
    r = open_reader(paths[0], paths[1], S, batch_blocks=4)
    assert isinstance(r, StreamReader)
    got0, got1, firsts = [], [], []
    try:
        for c0, c1, first, n in r:
            got0.append(c0.copy()); got1.append(c1.copy()); firsts.append((first, n))
    finally:
        # unblock the writer of the longer stream
        for p in paths:
            try:
                fd = os.open(p, os.O_RDONLY | os.O_NONBLOCK)
                while os.read(fd, 1 << 16):
                    pass
                os.close(fd)
            except OSError:
                pass
    for t in ths:
        t.join(timeout=10)
    assert firsts == [(0, 4), (4, 4), (8, 1)] and r.n_blocks == nb
    np.testing.assert_array_equal(np.concatenate(got0), a[:2 * S * nb])
    np.testing.assert_array_equal(np.concatenate(got1), b[:2 * S * nb])


def test_stream_reader_max_blocks_and_file_objects():
    import io
    from effex_b200.ingest import StreamReader
    S = 64
    a = np.arange(2 * S * 10, dtype=np.uint32).astype(np.uint8)
    r = StreamReader(io.BytesIO(a.tobytes()), io.BytesIO(a[::-1].copy().tobytes()), S, batch_blocks=3, max_blocks=7)
    out = [(first, n, c0.copy(), c1.copy()) for c0, c1, first, n in r]
    assert [(f, n) for f, n, _, _ in out] == [(0, 3), (3, 3), (6, 1)]
    np.testing.assert_array_equal(np.concatenate([o[2] for o in out]), a[:2 * S * 7])
    np.testing.assert_array_equal(np.concatenate([o[3] for o in out]), a[::-1][:2 * S * 7])


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
