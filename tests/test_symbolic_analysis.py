"""Host-side model compilation (capi.cu: symbolic_analysis) without a GPU: bionc_cuda_model_create runs the node
grouping, minimum-degree ordering, fill pattern, elimination-tree levels and the schedule choice on the host; only the
first launch touches the device.  The layouts of the golden models are pinned here (they decide shared-memory
residency and the joint-level schedule, i.e. performance), and random topologies must always compile."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
from helpers import model_path  # noqa: E402

KEYS = ("segments", "joint_rows", "nodes", "stored_blocks", "second_copies", "tree_levels", "row_storage_doubles",
        "zero_list", "systems_per_cta", "smem_bytes_per_cta", "schedule")


def layout_of(table):
    from bionc_b200.bionc_cuda import _lib

    lib = _lib.load()
    desc, keep = _lib.model_desc(table, with_markers=False)  # the marker table is the only device allocation of create
    handle = C.c_void_p()
    _lib.check(lib.bionc_cuda_model_create(C.byref(desc), 0, C.byref(handle)), "model_create")
    try:
        out = (C.c_int32 * 11)()
        _lib.check(lib.bionc_cuda_model_layout(handle, out), "model_layout")
        n, mj, m = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.bionc_cuda_model_counts(handle, C.byref(n), C.byref(mj), C.byref(m)), "model_counts")
        d = dict(zip(KEYS, [int(v) for v in out]))
        d["m"] = m.value
        return d
    finally:
        lib.bionc_cuda_model_destroy(handle)


EXPECTED = {
    # model: (nodes, stored blocks, second copies, tree levels, schedule)   schedule: 0 level-parallel, 1 serial, 2 two leaves + root
    "lower_limb": (3, 5, 3, 2, 2),
    "full_body": (9, 18, 9, 5, 0),
    "knee_feikes": (3, 6, 3, 3, 1),
    "pendulum_root": (2, 3, 0, 2, 1),
    "n_link_4": (7, 21, 11, 5, 0),
}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_layout_of_golden_models(name):
    from bionc_b200.bionc_cuda.model_table import ModelTable

    table = ModelTable.load(model_path(name))
    lay = layout_of(table)
    assert lay["segments"] == table.nb_segments and lay["joint_rows"] == table.nb_joint_constraints
    assert lay["m"] == table.nb_holonomic_constraints
    assert (lay["nodes"], lay["stored_blocks"], lay["second_copies"], lay["tree_levels"], lay["schedule"]) == EXPECTED[name]
    assert lay["systems_per_cta"] == 0 and lay["smem_bytes_per_cta"] == 0  # sized at the first launch, on the device


def test_random_topologies_compile_on_the_host():
    from test_gpu_random_models import CASES, random_table

    for seed, N, loops in CASES:
        table = random_table(1000 + seed, N, loops)
        lay = layout_of(table)
        rows = table.nb_joint_constraints
        assert lay["nodes"] * 3 >= rows and lay["nodes"] <= rows  # every node holds 1-3 real rows
        full = lay["nodes"] * (lay["nodes"] + 1) // 2
        assert lay["nodes"] <= lay["stored_blocks"] <= full or lay["nodes"] == 0
        assert 0 <= lay["second_copies"] <= lay["stored_blocks"]
        assert lay["tree_levels"] <= max(lay["nodes"], 0) and lay["schedule"] in (0, 1, 2)


def test_a_chain_of_three_spherical_joints_gets_the_two_leaf_schedule_only_with_two_free_ends():
    """hip-knee-ankle: the end joints are leaves of the knee.  With the pelvis pinned to the ground the hip joint node
    is coupled with the ground joint too and the shape is a chain of four nodes: no two-leaf schedule."""
    from bionc_b200.bionc_cuda.model_table import BLK_LIN3
    from test_gpu_random_models import assemble_table, random_table

    base = random_table(5, 4, 0)
    sph = (BLK_LIN3, np.concatenate([[0, 0, 1.0, 0], [0, 1.0, 0, 0], np.zeros(3)]))
    gsph = (BLK_LIN3, np.concatenate([np.zeros(4), [0, 1.0, 0, 0], np.zeros(3)]))
    args = (base.seg_phi_const, base.seg_geometry, base.seg_mass, base.seg_com, base.seg_J)
    free = assemble_table(*args, [(0, 1, [sph]), (1, 2, [sph]), (2, 3, [sph])])
    assert layout_of(free)["schedule"] == 2
    pinned = assemble_table(*args, [(-1, 0, [gsph]), (0, 1, [sph]), (1, 2, [sph]), (2, 3, [sph])])
    lay = layout_of(pinned)
    assert lay["nodes"] == 4 and lay["schedule"] in (0, 1)
