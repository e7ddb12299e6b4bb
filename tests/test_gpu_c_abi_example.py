"""The C ABI used from plain C (tests/c_abi/spherical_pendulum.c): compiled with gcc against include/bionc_cuda.h and
the in-tree library, run as its own process -- no Python, no torch on that side -- and its printed results compared
with the CPU oracle on the same model and states."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import REPO, rel_err  # noqa: E402


def _oracle_model():
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda.model_table import BLK_LIN3, BLK_NPARAM, ModelTable

    half_pi = 1.5707963267948966
    param = np.zeros((1, BLK_NPARAM))
    param[0, 5] = 1.0
    table = ModelTable(
        segment_names=["pendulum"], seg_phi_const=np.array([[np.cos(half_pi), np.cos(half_pi), 1.0, np.cos(half_pi)]]),
        seg_geometry=np.array([[1.0, half_pi, half_pi, half_pi]]), seg_mass=np.array([2.0]), seg_com=np.array([[0.0, -0.5, 0.0]]),
        seg_J=np.diag([0.1, 0.6, 0.1])[None], seg_rigid_row=np.array([3], dtype=np.int32), joint_names=["ground_spherical"],
        joint_kinds=["GroundJoint.Spherical"], blk_type=np.array([BLK_LIN3], dtype=np.int32), blk_parent=np.array([-1], dtype=np.int32),
        blk_child=np.array([0], dtype=np.int32), blk_joint=np.array([0], dtype=np.int32), blk_row_h=np.array([0], dtype=np.int32),
        blk_row_j=np.array([0], dtype=np.int32), blk_param=param,
    ).validate()
    return OracleModel(table)


def test_plain_c_program_matches_the_oracle(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    lib_dir = os.path.join(REPO, "bionc_b200", "lib")
    exe = str(tmp_path / "spherical_pendulum")
    subprocess.run(
        [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), os.path.join(REPO, "tests", "c_abi", "spherical_pendulum.c"),
         "-L", lib_dir, "-lbionc_cuda", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
    rows = {}
    for line in out.splitlines():
        f = line.split()
        if f[0] in ("counts", "launches"):
            rows[f[0]] = [int(x) for x in f[1:]]
        elif f[0] == "status":
            rows.setdefault("status", []).append(int(f[2]))
        else:
            rows[(f[0], int(f[1]))] = np.array([float(x) for x in f[2:]])
    assert rows["counts"] == [1, 3, 9] and rows["launches"] == [2] and rows["status"] == [0, 0, 0, 0]
    om = _oracle_model()
    t = np.arange(101) * 0.01
    for s in range(2):
        Q0, Qd0 = rows[("Q0", s)], rows[("Qd0", s)]
        assert np.abs(om.holonomic_constraints(Q0)).max() < 1e-15  # the C program's states are consistent
        qdd, lam = om.forward_dynamics(Q0, Qd0)
        assert rel_err(rows[("Qdd", s)], qdd) < 1e-10 and rel_err(rows[("lam", s)], lam) < 1e-10
        Qf, Qdf = om.rk4(Q0, Qd0, t, normalize=True)
        assert rel_err(np.concatenate([rows[("Qf", s)], rows[("Qdf", s)]]), np.concatenate([Qf, Qdf])) < 1e-8
    # gravity acts along -z (biomechanical_model.py:319-333): the segment released along -y swings about x, u stays put
    assert np.abs(rows[("Qdf", 0)][:3]).max() < 1e-13 and abs(rows[("Qdf", 0)][10]) > 0.1
