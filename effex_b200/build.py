"""Build libeffex_fx.so in-tree with nvcc for sm_100a (the only target).

    python -m effex_b200.build            # build if sources are newer than the .so
    python -m effex_b200.build --force -v # rebuild, print ptxas resource usage
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeffex_fx.so")
SOURCES = ["fx_abi.cu", "fx_csv.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libeffex_fx.so cannot be built")
    return exe


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "effex_fx.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """`defines`/`out` build an experiment variant (tools/kbench.py); the product is the default.
    Safe against concurrent callers (the ranks of a torchrun launch): one builds under an exclusive lock into a
    temporary file that is renamed into place, the others wait and find a fresh library."""
    if out is None and not force and not _stale():
        return LIB
    out = out or LIB
    import fcntl
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out == LIB and not force and not _stale():        # somebody else built it while we waited
                return LIB
            tmp = f"{out}.tmp{os.getpid()}"
            cmd = [_nvcc(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-pthread", "-shared",
                   "-o", tmp] + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libeffex_fx.so")
            os.replace(tmp, out)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
