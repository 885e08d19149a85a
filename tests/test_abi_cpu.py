"""CPU-side checks of the boundary: the library builds/loads, exports every
symbol include/effex_fx.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from effex_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "effex_fx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = _declared()
    assert len(names) >= 25
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.fx_abi_version() == _lib.FX_ABI_VERSION == 2


def test_config_struct_layout():
    # must match `fx_config` in include/effex_fx.h (int32 x4, int64, int32 x2)
    assert C.sizeof(_lib.FxConfig) == 32
    assert _lib.FxConfig.num_samp.offset == 16


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    bad = _lib.FxConfig(0, 33, 4096, 1, 2**18, 1, 0)       # ntaps > 32, like cuSignal's cap
    assert lib.fx_create(C.byref(bad), C.byref(h)) == _lib.FX_ERR_UNSUPPORTED
    for nb in (7, 65537):                                  # nbins outside [8, 65536] (any integer inside is fine)
        bad = _lib.FxConfig(0, 4, nb, 1, 2**18, 1, 0)
        assert lib.fx_create(C.byref(bad), C.byref(h)) == _lib.FX_ERR_UNSUPPORTED
    bad = _lib.FxConfig(0, 4, 4096, 1, 100, 1, 0)          # less than one frame
    assert lib.fx_create(C.byref(bad), C.byref(h)) == _lib.FX_ERR_INVALID
    assert b"frame" in lib.fx_last_error(None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.FxConfig(0, 4, 4096, 1, 2**18, 1, 0)
    assert lib.fx_create(C.byref(cfg), C.byref(h)) == _lib.FX_ERR_CUDA
    assert b"no CPU fallback" in lib.fx_last_error(None)
    from effex_b200.engine import FxEngine, FxError
    with pytest.raises(FxError):
        FxEngine(2**18, 4096)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "effex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle-", ""), f


def test_rot_vector_matches_reference_expression():
    from effex_b200.engine import rot_vector
    from oracle import fx_oracle as orc
    for tau in (0.0, 37 / 2.4e6, -5.2e-6, 1.234e-3):
        np.testing.assert_allclose(rot_vector(4096, 2.4e6, 1.4204e9, tau),
                                   orc.rot_vector(4096, 2.4e6, 1.4204e9, tau), atol=2e-6)
