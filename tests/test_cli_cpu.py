"""The CLI keeps the reference's flags, dests, defaults and types (effex.py:703-770)."""
from effex_b200.cli import build_parser


def test_flags_match_reference():
    p = build_parser()
    d = p.parse_args([])
    assert (d.run_time, d.bandwidth, d.fc, d.num_samp, d.nfft, d.gain, d.mode, d.omit_plot, d.loglevel) == (
        1, 2.4e6, 1.4204e9, 2**18, 2**12, 49.6, 'spectrum', False, 'INFO')
    a = p.parse_args("-T 60 -B 2.4e6 -F 1.4204e9 -N 262144 -R 4096 -G 29.7 -M spectrum".split())   # run.sh:5
    assert a.run_time == 60.0 and a.gain == 29.7 and a.nfft == 4096
    a = p.parse_args(["--omit_plot", "False"])
    assert a.omit_plot is True          # quirk Q7: type=bool, any non-empty string is truthy
    for bad in (["--mode", "foo"], ["--loglevel", "TRACE"]):
        try:
            p.parse_args(bad)
        except SystemExit:
            continue
        raise AssertionError(bad)


def test_lazy_tiled_recording_is_the_tiled_recording():
    import numpy as np
    from effex_b200 import synth
    S = 1024
    a0, a1 = synth.tiled_recording(37, S, base_blocks=8, delay=3)
    b0, b1 = synth.tiled_recording_lazy(37, S, base_blocks=8, delay=3, window_blocks=5)
    assert b0.size == a0.size and len(b1) == a1.size
    for lo, hi in ((0, 2 * S), (2 * S * 3, 2 * S * 8), (2 * S * 30, 2 * S * 35), (2 * S * 36, 2 * S * 37), (2 * S * 7, 2 * S * 12)):
        assert np.array_equal(a0[lo:hi], b0[lo:hi]) and np.array_equal(a1[lo:hi], b1[lo:hi])
    import pytest
    with pytest.raises(ValueError):
        b0[0:2 * S * 37]                       # longer than the window
