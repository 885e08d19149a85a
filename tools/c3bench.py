#!/usr/bin/env python
"""A/B of library variants on C3 (nbins 65536, num_samp 2^24, 2 block pairs): step time and per-kernel
times (CUDA events around every launch are not available through the ABI, so the per-kernel split comes
from one ncu duration pass per variant).  Run under gpurun.
usage: python tools/c3bench.py variant.so [variant2.so ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, torch
sys.path.insert(0, %r)
from effex_b200 import synth
from effex_b200.engine import FxEngine
S, N, nb = 2**24, 65536, 2
raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=2)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=nb)
out = (torch.empty((nb, N), dtype=torch.complex64, device="cuda"), None, None)
reps = int(os.environ.get("C3_REPS", "30"))
for _ in range(3): eng.process(d0, d1, nb, out=out)
eng.sync()
t0 = time.perf_counter()
for _ in range(reps): eng.process(d0, d1, nb, out=out)
eng.sync()
us = (time.perf_counter() - t0) / reps * 1e6
print("%%-24s C3 pass %%7.1f us  %%8.0f Msamples/s  checksum %%.6e" %% (os.path.basename(os.environ.get("EFFEX_FX_LIB", "default")), us, nb * S / us, float(out[0].abs().sum().item())))
''' % ROOT

for lib in sys.argv[1:]:
    env = dict(os.environ, EFFEX_FX_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
    if os.environ.get("C3_NCU", "1") == "1":
        env["C3_REPS"] = "2"
        csv = os.path.join(ROOT, "gpurun_out", "c3_" + os.path.basename(lib) + ".csv")
        subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", csv,
                        sys.executable, "-c", CHILD], env=env, check=False, stdout=subprocess.DEVNULL)
        import csv as _csv
        rows = [r for r in _csv.reader(open(csv)) if len(r) > 5]
        hdr = next(r for r in rows if "Kernel Name" in r)
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        body = rows[rows.index(hdr) + 2:]
        last = body[-4 * 1:] if False else body
        # the last pass's launches
        names = [(r[ki].split("(")[0][-40:], float(r[vi].replace(",", ""))) for r in body]
        n_per = len(names) // 5 if len(names) % 5 == 0 else 0
        tail = names[-(n_per or 6):]
        print("    last pass: " + "  ".join("%s %.1f us" % (n.split("::")[-1], v / 1e3) for n, v in tail))
