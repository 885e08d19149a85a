"""Random shapes: every fast path (fused 256..4096 bins, head/tail 8192..65536 bins, ragged num_samp,
partial super-frames, with and without DC removal) against the unfused kernels (tools/fuzz_shapes.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [7, 8])
def test_fast_paths_agree_with_unfused_kernels_on_random_shapes(seed):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_shapes.py"), "16", str(seed)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "fuzz ok: 16 cases" in res.stdout


@pytest.mark.gpu
def test_lag_search_every_transform_length_against_unfused_passes():
    """M = G*4096 for G = 2 ... 256: every split of the register lag head, ragged n, several blocks (tools/fuzz_lag.py)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_lag.py"), "5"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "lag fuzz ok: 15 sizes" in res.stdout
