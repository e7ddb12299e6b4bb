"""In-tree build of the CUDA library: ``nvcc`` cross-compiles for sm_100a (no GPU needed)."""

from __future__ import annotations

import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "csrc", "capi.cu")
DEPS = [os.path.join(ROOT, "csrc", f) for f in ("capi.cu", "kernels.cuh", "dynamics.cuh", "model.cuh", "tmem.cuh")] + [
    os.path.join(os.path.dirname(ROOT), "include", "bionc_cuda.h")
]
# BIONC_LIB_SUFFIX builds an experimental variant next to the library (profiles/: occupancy sweeps); _lib.py loads
# $BIONC_CUDA_LIB instead of the default when it is set
OUT = os.path.join(ROOT, "lib", "libbionc_cuda%s.so" % os.environ.get("BIONC_LIB_SUFFIX", ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/capi.cu`` into ``lib/libbionc_cuda.so`` if sources are newer than the library."""
    if not force and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = os.environ.get("BIONC_NVCC_EXTRA", "").split()
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra] + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
