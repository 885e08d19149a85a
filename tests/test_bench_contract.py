"""bench.py's JSON contract: the reference arm runs on CPU cores alone; the product arm needs the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args, env=None):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=900, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the oracle port on the host cores, same metric/config keys, no GPU touched."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-blocks", "1", env=env)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "fx_msamples_per_s_per_channel_pair" and d["unit"] == "Msamples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"]
    assert set(d["config"]) == {"workload", "parallelism", "l2"}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.gpu
def test_product_arm_line():
    d = run_bench("--steps", "20", "--warmup", "3", "--cpu-blocks", "1")
    assert BASE_KEYS | {"gpu_launches", "roofline", "clocks", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["dtype"] == "f32"
    assert d["value"] > 50_000                      # Msamples/s: orders of magnitude above any CPU path
    assert d["gpu_launches"] >= 3 * 20              # byte sums + fused + finalize per step
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 4 * 262144 * 550 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "fused_kernel_stag" in r["kernel"]
    assert r["fp32_pipe"] is None or 0 < r["fp32_pipe"]["frac"] < 1
    assert d["e2e_rows_match_device_rows"] is True
    assert set(d["config"]) == {"workload", "parallelism", "l2"}          # identical in both arms
    assert 0 < e["frac_of_copy_only"] <= 1.05 and e["copy_only"]["value"] > 0
    c = d["configs"]
    assert set(c) == {"c2", "c3", "c4", "c5"} and not any("error" in v for v in c.values()), c
    assert c["c2"]["integer_lag"] == 37
    assert c["c4"]["frames_integrated"] == c["c4"]["frames_expected"]
    assert c["c5"]["csv_format"]["bytes"] > 0 and c["c3"]["value"] > 20_000
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
