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
