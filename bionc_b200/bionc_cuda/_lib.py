"""ctypes binding of the C ABI in ``include/bionc_cuda.h`` (``bionc_b200/lib/libbionc_cuda.so``).

There is deliberately no fallback: if the shared library is missing or a call fails, the backend
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""

from __future__ import annotations

import ctypes as C
import os

LIB_PATH = os.environ.get("BIONC_CUDA_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libbionc_cuda.so")

# every symbol include/bionc_cuda.h declares (tests/test_c_abi.py checks the header against this list)
SYMBOLS = (
    "bionc_cuda_model_create",
    "bionc_cuda_model_destroy",
    "bionc_cuda_model_counts",
    "bionc_cuda_model_layout",
    "bionc_cuda_model_joint_rows",
    "bionc_cuda_model_specialize",
    "bionc_cuda_model_spec_source",
    "bionc_cuda_model_specialized",
    "bionc_cuda_forward_dynamics",
    "bionc_cuda_dense_workspace_bytes",
    "bionc_cuda_forward_dynamics_dense",
    "bionc_cuda_rk4",
    "bionc_cuda_eval",
    "bionc_cuda_mass_matrix",
    "bionc_cuda_forward_dynamics_host",
    "bionc_cuda_rk4_host",
    "bionc_cuda_measure_fp64_peak",
    "bionc_cuda_debug_poison",
    "bionc_cuda_launch_count",
    "bionc_cuda_last_error",
    "bionc_cuda_version",
)

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class BioncModelDesc(C.Structure):
    _fields_ = [
        ("nb_segments", C.c_int32),
        ("nb_blocks", C.c_int32),
        ("seg_phi_const", c_double_p),
        ("seg_mass", c_double_p),
        ("seg_com", c_double_p),
        ("seg_J", c_double_p),
        ("seg_rigid_row", c_int32_p),
        ("blk_type", c_int32_p),
        ("blk_parent", c_int32_p),
        ("blk_child", c_int32_p),
        ("blk_row_h", c_int32_p),
        ("blk_row_j", c_int32_p),
        ("blk_param", c_double_p),
        ("nb_markers", C.c_int32),
        ("marker_segment", c_int32_p),
        ("marker_position", c_double_p),
    ]


class BioncFextDesc(C.Structure):
    _fields_ = [
        ("nb_forces", C.c_int32),
        ("segment", c_int32_p),
        ("kind", c_int32_p),
        ("param", c_double_p),
        ("values", C.c_void_p),  # optional per-system (torque, force): (values_steps, B, F, 6)
        ("values_steps", C.c_int64),
    ]


class BioncDynOptions(C.Structure):
    _fields_ = [
        ("use_stabilization", C.c_int32),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("fext", C.POINTER(BioncFextDesc)),
        ("inertia", C.c_void_p),
    ]


def model_desc(table, with_markers: bool = True):
    """``BioncModelDesc`` for a (validated) ModelTable plus the arrays it points into (keep them alive as long as
    the descriptor is used)."""
    import numpy as np

    def ptr(a, typ):
        return a.ctypes.data_as(typ)

    k = dict(
        phic=np.ascontiguousarray(table.seg_phi_const, dtype=np.float64),
        mass=np.ascontiguousarray(table.seg_mass, dtype=np.float64),
        com=np.ascontiguousarray(table.seg_com, dtype=np.float64),
        J=np.ascontiguousarray(table.seg_J, dtype=np.float64).reshape(-1, 9),
        rrow=np.ascontiguousarray(table.seg_rigid_row, dtype=np.int32),
        btype=np.ascontiguousarray(table.blk_type, dtype=np.int32),
        bpar=np.ascontiguousarray(table.blk_parent, dtype=np.int32),
        bchi=np.ascontiguousarray(table.blk_child, dtype=np.int32),
        browh=np.ascontiguousarray(table.blk_row_h, dtype=np.int32),
        browj=np.ascontiguousarray(table.blk_row_j, dtype=np.int32),
        bpar_d=np.ascontiguousarray(table.blk_param, dtype=np.float64),
        mkseg=np.ascontiguousarray(table.marker_segment, dtype=np.int32),
        mkpos=np.ascontiguousarray(table.marker_position, dtype=np.float64),
    )
    desc = BioncModelDesc(
        table.nb_segments, table.nb_blocks,
        ptr(k["phic"], c_double_p), ptr(k["mass"], c_double_p), ptr(k["com"], c_double_p), ptr(k["J"], c_double_p),
        ptr(k["rrow"], c_int32_p), ptr(k["btype"], c_int32_p), ptr(k["bpar"], c_int32_p), ptr(k["bchi"], c_int32_p),
        ptr(k["browh"], c_int32_p), ptr(k["browj"], c_int32_p), ptr(k["bpar_d"], c_double_p),
        table.nb_markers if with_markers else 0, ptr(k["mkseg"], c_int32_p), ptr(k["mkpos"], c_double_p),
    )
    return desc, k


class BioncCudaError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built -- no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BioncCudaError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`); bionc_cuda has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    optp = C.POINTER(BioncDynOptions)
    lib.bionc_cuda_model_create.argtypes = [C.POINTER(BioncModelDesc), C.c_int, C.POINTER(vp)]
    lib.bionc_cuda_model_destroy.argtypes = [vp]
    lib.bionc_cuda_model_counts.argtypes = [vp, c_int32_p, c_int32_p, c_int32_p]
    lib.bionc_cuda_model_layout.argtypes = [vp, c_int32_p]
    lib.bionc_cuda_model_joint_rows.argtypes = [vp, c_int32_p]
    lib.bionc_cuda_model_specialize.argtypes = [vp, optp, i32, i32]
    lib.bionc_cuda_model_spec_source.argtypes = [vp, optp, i32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.bionc_cuda_model_specialized.argtypes = [vp, C.POINTER(i64)]
    lib.bionc_cuda_forward_dynamics.argtypes = [vp, vp, vp, optp, vp, vp, vp, i64, vp]
    lib.bionc_cuda_dense_workspace_bytes.argtypes = [vp, i64]
    lib.bionc_cuda_dense_workspace_bytes.restype = C.c_size_t
    lib.bionc_cuda_forward_dynamics_dense.argtypes = [vp, vp, vp, optp, vp, i64, vp, vp, vp, vp, C.c_size_t, vp]
    lib.bionc_cuda_rk4.argtypes = [vp, vp, vp, dbl, vp, i64, i32, optp, vp, i64, vp, vp, i64, vp]
    lib.bionc_cuda_eval.argtypes = [vp, i32, vp, optp, vp, i64, vp]
    lib.bionc_cuda_mass_matrix.argtypes = [vp, vp]
    lib.bionc_cuda_forward_dynamics_host.argtypes = [vp, vp, vp, optp, vp, vp, vp, i64]
    lib.bionc_cuda_rk4_host.argtypes = [vp, vp, vp, dbl, vp, i64, i32, optp, vp, i64]
    lib.bionc_cuda_measure_fp64_peak.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p]
    lib.bionc_cuda_debug_poison.argtypes = [i32]
    lib.bionc_cuda_launch_count.restype = C.c_int64
    lib.bionc_cuda_last_error.restype = C.c_char_p
    lib.bionc_cuda_version.restype = C.c_char_p
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("bionc_cuda_launch_count", "bionc_cuda_last_error", "bionc_cuda_version", "bionc_cuda_dense_workspace_bytes"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().bionc_cuda_last_error().decode(errors="replace")
        raise BioncCudaError(f"{what} failed with code {rc}: {msg}")
