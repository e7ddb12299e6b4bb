"""The C-ABI shared library loads without a GPU and exports every symbol include/bionc_cuda.h declares.
No compute call is made here (there is no device in the build container)."""
import ctypes
import os
import re

import pytest

from helpers import REPO

HEADER = os.path.join(REPO, "include", "bionc_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bionc_cuda_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("bionc_cuda_model_create", "bionc_cuda_forward_dynamics", "bionc_cuda_rk4", "bionc_cuda_eval",
                 "bionc_cuda_rk4_host", "bionc_cuda_forward_dynamics_host", "bionc_cuda_mass_matrix"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from bionc_b200 import build
    from bionc_b200.bionc_cuda import _lib

    build.build()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == declared_symbols(), "python binding and header disagree"
    lib.bionc_cuda_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.bionc_cuda_version()


def test_library_is_native_sm100a_code():
    """The shipped .so must contain sm_100a SASS (not PTX-only, not another arch)."""
    import shutil
    import subprocess

    from bionc_b200.bionc_cuda import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_invalid_arguments_return_error_codes_without_a_device():
    from bionc_b200.bionc_cuda import _lib

    lib = _lib.load()
    assert lib.bionc_cuda_model_create(None, 0, None) == -1
    assert b"null" in lib.bionc_cuda_last_error()
    desc = _lib.BioncModelDesc()
    desc.nb_segments = 40  # more than one lane per segment allows
    handle = ctypes.c_void_p()
    assert lib.bionc_cuda_model_create(ctypes.byref(desc), 0, ctypes.byref(handle)) == -1


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: constructing a model without a CUDA device raises."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from bionc_b200.bionc_cuda import BiomechanicalModel, BioncCudaError
    from helpers import model_path

    with pytest.raises(BioncCudaError):
        BiomechanicalModel.load(model_path("lower_limb"))


def test_struct_layouts_and_constants_match_the_header(tmp_path):
    """The ctypes mirrors in _lib.py against the C compiler's view of include/bionc_cuda.h (sizes, field offsets) and
    the Python-side constants against the header's #defines."""
    import shutil
    import subprocess

    from bionc_b200.bionc_cuda import _lib, model_table as mt

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    fields = {
        "BioncModelDesc": [n for n, _ in _lib.BioncModelDesc._fields_],
        "BioncFextDesc": [n for n, _ in _lib.BioncFextDesc._fields_],
        "BioncDynOptions": [n for n, _ in _lib.BioncDynOptions._fields_],
    }
    consts = ["BIONC_BLK_LIN3", "BIONC_BLK_DOT", "BIONC_BLK_CLEN", "BIONC_BLK_SOP", "BIONC_BLK_NPARAM", "BIONC_FEXT_NPARAM",
              "BIONC_FEXT_GLOBAL_ON_PROXIMAL", "BIONC_FEXT_GLOBAL_LOCAL_POINT", "BIONC_FEXT_GLOBAL", "BIONC_FEXT_LOCAL",
              "BIONC_EVAL_PHI_R", "BIONC_EVAL_K_R", "BIONC_EVAL_PHI_J", "BIONC_EVAL_K_J", "BIONC_EVAL_PHI_H", "BIONC_EVAL_K_H",
              "BIONC_EVAL_BIAS_H", "BIONC_EVAL_FEXT", "BIONC_EVAL_COM", "BIONC_EVAL_MARKERS", "BIONC_EVAL_KINETIC",
              "BIONC_EVAL_POTENTIAL", "BIONC_EVAL_RESIDUALS", "BIONC_STATUS_SINGULAR", "BIONC_STATUS_NONFINITE",
              "BIONC_STATUS_SMALL_PIVOT", "BIONC_STATUS_PIVOTED"]
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "bionc_cuda.h"', "int main(void) {"]
    for st, names in fields.items():
        lines.append(f'printf("sizeof {st} %zu\\n", sizeof({st}));')
        lines += [f'printf("offset {st}.{n} %zu\\n", offsetof({st}, {n}));' for n in names]
    lines += [f'printf("const {c} %d\\n", (int){c});' for c in consts]
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        kind, name, val = ln.split()
        got[(kind, name)] = int(val)
    for st in fields:
        cls = getattr(_lib, st)
        assert got[("sizeof", st)] == ctypes.sizeof(cls), st
        for n in fields[st]:
            assert got[("offset", f"{st}.{n}")] == getattr(cls, n).offset, f"{st}.{n}"
    c = {name: val for (kind, name), val in got.items() if kind == "const"}
    assert (c["BIONC_BLK_LIN3"], c["BIONC_BLK_DOT"], c["BIONC_BLK_CLEN"], c["BIONC_BLK_SOP"]) == (mt.BLK_LIN3, mt.BLK_DOT, mt.BLK_CLEN, mt.BLK_SOP)
    assert c["BIONC_BLK_NPARAM"] == mt.BLK_NPARAM and c["BIONC_FEXT_NPARAM"] == mt.FEXT_NPARAM
    assert (c["BIONC_FEXT_GLOBAL_ON_PROXIMAL"], c["BIONC_FEXT_GLOBAL_LOCAL_POINT"], c["BIONC_FEXT_GLOBAL"], c["BIONC_FEXT_LOCAL"]) == (
        mt.FEXT_GLOBAL_ON_PROXIMAL, mt.FEXT_GLOBAL_LOCAL_POINT, mt.FEXT_GLOBAL, mt.FEXT_LOCAL)
    evals = ["PHI_R", "K_R", "PHI_J", "K_J", "PHI_H", "K_H", "BIAS_H", "FEXT", "COM", "MARKERS", "KINETIC", "POTENTIAL", "RESIDUALS"]
    import importlib

    try:
        bm_mod = importlib.import_module("bionc_b200.bionc_cuda.biomechanical_model")
    except Exception:  # torch missing: the constants are plain ints at module level, nothing else to check
        bm_mod = None
    if bm_mod is not None:
        for k, e in enumerate(evals):
            assert c["BIONC_EVAL_" + e] == k == getattr(bm_mod, "EVAL_" + e)
    assert (c["BIONC_STATUS_SINGULAR"], c["BIONC_STATUS_NONFINITE"], c["BIONC_STATUS_SMALL_PIVOT"], c["BIONC_STATUS_PIVOTED"]) == (1, 2, 4, 8)
