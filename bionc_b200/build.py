"""In-tree build of the CUDA library: ``nvcc`` cross-compiles for sm_100a (no GPU needed)."""

from __future__ import annotations

import os
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.abspath(__file__))
# translation units: the C ABI with the evaluator kernels, and the dynamics kernels' instantiations (csrc/instances.cuh),
# compiled in parallel and linked into one library
UNITS = ("capi", "inst_fd", "inst_rk4_lean", "inst_rk4_general_nofx", "inst_rk4_general_fx")
HEADERS = ("kernels.cuh", "dynamics.cuh", "model.cuh", "tmem.cuh", "instances.cuh")
DEPS = [os.path.join(ROOT, "csrc", f) for f in tuple(u + ".cu" for u in UNITS) + HEADERS] + [
    os.path.join(os.path.dirname(ROOT), "include", "bionc_cuda.h")
]
# BIONC_LIB_SUFFIX builds an experimental variant next to the library (profiles/: occupancy sweeps); _lib.py loads
# $BIONC_CUDA_LIB instead of the default when it is set
OUT = os.path.join(ROOT, "lib", "libbionc_cuda%s.so" % os.environ.get("BIONC_LIB_SUFFIX", ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
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
    """Compile ``csrc/*.cu`` into ``lib/libbionc_cuda.so`` if sources are newer than the library."""
    if not force and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = os.environ.get("BIONC_NVCC_EXTRA", "").split()
    flags = [*NVCC_FLAGS, *extra] + (["-Xptxas", "-v"] if verbose else [])
    with tempfile.TemporaryDirectory(prefix="bionc_build_") as tmp:
        objs = [os.path.join(tmp, u + ".o") for u in UNITS]
        cmds = [[nvcc_path(), *flags, "-c", "-o", o, os.path.join(ROOT, "csrc", u + ".cu")] for u, o in zip(UNITS, objs)]
        procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
        logs = [p.communicate()[0] for p in procs]
        for c, p, log in zip(cmds, procs, logs):
            if p.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(c) + "\n" + log)
        link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT + ".tmp", *objs]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
        os.replace(OUT + ".tmp", OUT)
    if verbose:
        print("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
