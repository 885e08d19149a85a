"""The index algebra of the kernels, emulated in numpy and checked against numpy.fft on the CPU:
tools/proto_fft4096.py (16x16x16 dataflow) and tools/proto_superframe.py (nbins below and above 4096)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_superframe_and_bigfft_index_algebra():
    assert _load("proto_superframe").check(seed=3) < 1e-12


def test_fft4096_dataflow_emulation():
    import numpy as np
    proto = _load("proto_fft4096")
    rng = np.random.default_rng(4)
    w = rng.standard_normal(4096) + 1j * rng.standard_normal(4096)
    got = proto.fft4096_emulated(w)
    ref = np.fft.fft(w)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
