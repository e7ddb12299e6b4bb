"""Model-specialised kernels without a GPU: the generated translation unit carries the model's tables, names the kernel
variant of the call shape, and NVRTC compiles it for sm_100a (NVRTC needs no device; skipped where libnvrtc is absent)."""
import ctypes as C
import ctypes.util
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
from helpers import model_path  # noqa: E402


def _handle(name):
    from bionc_b200.bionc_cuda import _lib
    from bionc_b200.bionc_cuda.model_table import ModelTable

    lib = _lib.load()
    table = ModelTable.load(model_path(name))
    desc, keep = _lib.model_desc(table, with_markers=False)
    h = C.c_void_p()
    _lib.check(lib.bionc_cuda_model_create(C.byref(desc), 0, C.byref(h)), "model_create")
    return lib, h, table


def _source(lib, h, opt, spc=32):
    from bionc_b200.bionc_cuda import _lib

    need = C.c_size_t()
    _lib.check(lib.bionc_cuda_model_spec_source(h, C.byref(opt), spc, None, 0, C.byref(need)), "spec_source")
    buf = C.create_string_buffer(need.value)
    _lib.check(lib.bionc_cuda_model_spec_source(h, C.byref(opt), spc, buf, need.value, None), "spec_source")
    return buf.value.decode()


def _forces(lib_mod):
    seg, kind, par = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(24)
    fd = lib_mod.BioncFextDesc(1, seg.ctypes.data_as(lib_mod.c_int32_p), kind.ctypes.data_as(lib_mod.c_int32_p), par.ctypes.data_as(lib_mod.c_double_p))
    return fd, (seg, kind, par)


def test_source_names_the_variant_of_the_call_shape():
    from bionc_b200.bionc_cuda import _lib

    lib, h, table = _handle("lower_limb")
    try:
        opt = _lib.BioncDynOptions()
        src = _source(lib, h, opt)
        assert "constexpr int kSegments = 4;" in src and '#include "kernels.cuh"' in src
        assert "bionc::rk4_kernel<false, false, 128, 32, 0, false>" in src  # lean path, 4 warps, 32 systems per CTA
        assert src.count("0x") > 4 * 30  # the segments' constants as hexadecimal floating literals (exact)
        opt.use_stabilization = 1
        assert "bionc::rk4_kernel<false, true, 128, 32, 0, false>" in _source(lib, h, opt)
        fd, keep = _forces(_lib)
        opt.fext = C.pointer(fd)
        assert "bionc::rk4_kernel<true, true, 128, 32, 0, true>" in _source(lib, h, opt)  # forces: the general path
        opt.inertia = 1  # only the presence is read
        assert "bionc::rk4_kernel<true, true, 128, 16, 1, true>" in _source(lib, h, opt, spc=16)
        assert lib.bionc_cuda_model_spec_source(h, C.byref(opt), 24, None, 0, None) != 0  # 16 or 32 systems per CTA
    finally:
        lib.bionc_cuda_model_destroy(h)


@pytest.mark.parametrize("name,spc,forces", [("pendulum_force_batch", 32, True), ("knee_feikes", 32, False), ("n_link_4", 16, False), ("lower_limb", 32, False)])
def test_nvrtc_compiles_the_specialised_kernel_without_a_gpu(name, spc, forces):
    from bionc_b200.bionc_cuda import _lib

    if not (ctypes.util.find_library("nvrtc") or os.path.exists("/usr/local/cuda/lib64/libnvrtc.so")):
        pytest.skip("libnvrtc not installed")
    lib, h, table = _handle(name)
    try:
        opt = _lib.BioncDynOptions()
        if forces:
            fd, keep = _forces(_lib)
            opt.fext = C.pointer(fd)
        cubin = C.c_int64()
        assert lib.bionc_cuda_model_specialized(h, C.byref(cubin)) == 0 and cubin.value == 0
        _lib.check(lib.bionc_cuda_model_specialize(h, C.byref(opt), spc, 0), "model_specialize (compile only)")
        assert lib.bionc_cuda_model_specialized(h, C.byref(cubin)) == 0  # compiled, not loaded: nothing to launch
        assert cubin.value > 10_000
    finally:
        lib.bionc_cuda_model_destroy(h)


def test_build_decisions_follow_the_model():
    """Serial joint-level schedule -> compact working set; more than four joint-level nodes -> the factorisation keeps its
    loops (and the blob in shared memory, so no compact layout); the register budget follows the residency on a B200."""
    from bionc_b200.bionc_cuda import _lib

    expect = {
        "knee_feikes": (32, ["#define BIONC_SPEC_COMPACT 1", "#define BIONC_SPEC_MINCTAS 4"], ["BIONC_SPEC_ROLL_SOLVE"]),
        "pendulum_root": (32, ["#define BIONC_SPEC_COMPACT 1", "#define BIONC_SPEC_MINCTAS 8"], ["BIONC_SPEC_ROLL_SOLVE"]),
        "n_link_4": (16, ["#define BIONC_SPEC_ROLL_SOLVE 1", "#define BIONC_SPEC_MINCTAS 3"], ["BIONC_SPEC_COMPACT"]),
        "lower_limb": (32, ["#define BIONC_SPEC_MINCTAS 4"], ["BIONC_SPEC_COMPACT", "BIONC_SPEC_ROLL_SOLVE"]),
    }
    for name, (spc, present, absent) in expect.items():
        lib, h, table = _handle(name)
        try:
            head = _source(lib, h, _lib.BioncDynOptions(), spc).split('#include "model.cuh"')[0]
            for text in present:
                assert text in head, (name, text, head)
            for text in absent:
                assert text not in head, (name, text, head)
        finally:
            lib.bionc_cuda_model_destroy(h)
