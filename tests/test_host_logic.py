"""Host-side logic that needs no GPU: the model table, the workload generators, the FLOP counts,
the generic RK4 mirror and the external-force lowering."""
import os

import numpy as np
import pytest

from bionc_oracle import OracleModel
from conftest import golden_case_names
from helpers import GOLDEN, load_case, model_path

from bionc_b200.bionc_cuda import ExternalForceSet, ModelTable
from bionc_b200.bionc_cuda.model_table import BLK_ROWS, m4_from_natural_parameters
from bionc_b200.utils import roofline, states


@pytest.mark.parametrize("name", golden_case_names())
def test_model_table_roundtrip_and_counts(name, tmp_path):
    t = ModelTable.load(model_path(name))
    case = load_case(name)
    assert t.nb_holonomic_constraints == case["ref_phi_h"].shape[1]
    assert t.nb_joint_constraints == case["ref_phi_j"].shape[1]
    assert 12 * t.nb_segments == case["Q"].shape[1]
    np.testing.assert_allclose(t.mass_matrix, case["ref_mass_matrix"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(t.seg_gravity.reshape(-1), case["ref_gravity"], rtol=0, atol=1e-13)
    f = tmp_path / "t.npz"
    t.save(str(f))
    t2 = ModelTable.load(str(f))
    for k in ModelTable._ARRAYS:
        np.testing.assert_array_equal(getattr(t, k), getattr(t2, k))
    assert t2.segment_names == t.segment_names and t2.joint_kinds == t.joint_kinds


def test_holonomic_row_order_of_the_knee():
    # HIP 0-2 | rigid THIGH 3-8 | medial 9 | lateral 10 | ACL 11 | PCL 12 | MCL 13 | rigid SHANK 14-19
    t = ModelTable.load(model_path("knee_feikes"))
    assert list(t.seg_rigid_row) == [3, 14]
    assert list(t.blk_row_h) == [0, 9, 10, 11, 12, 13]
    assert [BLK_ROWS[int(x)] for x in t.blk_type] == [3, 1, 1, 1, 1, 1]


def test_cartesian_to_natural_inertia_matches_reference_models():
    """The as-coded conversion (natural_inertial_parameters.py:298-352) against values the reference itself
    produced for the golden models (orthogonal box, skewed bbox of tests/test_joints.py, lower-limb segments)."""
    from bionc_b200.bionc_cuda.model_table import natural_inertial_parameters_from_cartesian as conv

    t = ModelTable.load(model_path("tj_revolute"))
    c, J = conv(1.0, [0, 0, 0], np.eye(3), 1.0, np.pi / 2, np.pi / 2, np.pi / 2)
    np.testing.assert_allclose(c, t.seg_com[0], atol=1e-14)
    np.testing.assert_allclose(np.triu(J), np.triu(t.seg_J[0]), atol=1e-14)
    c, J = conv(1.1, [0.1, 0.11, 0.111], np.diag([1.1, 1.2, 1.3]), 1.5, np.pi / 1.9, np.pi / 2.3, np.pi / 2.1)
    np.testing.assert_allclose(c, t.seg_com[1], atol=1e-14)
    np.testing.assert_allclose(np.triu(J), np.triu(t.seg_J[1]), atol=1e-13)
    ll = ModelTable.load(model_path("lower_limb"))
    inertia = [(10.65, (0, -0.05, 0), (0.10, 0.11, 0.09)), (9.2, (0, -0.17, 0), (0.13, 0.05, 0.13)),
               (3.6, (0, -0.10, 0), (0.040, 0.020, 0.040)), (0.9, (0.02, -0.03, 0), (0.0040, 0.0045, 0.0045))]
    for i, (m, com, I) in enumerate(inertia):
        L, al, be, ga = ll.seg_geometry[i]
        c, J = conv(m, com, np.diag(I), L, al, be, ga)
        np.testing.assert_allclose(c, ll.seg_com[i], atol=1e-13)
        np.testing.assert_allclose(np.triu(J), np.triu(ll.seg_J[i]), atol=1e-13)


def test_m4_matches_reference_formula():
    rng = np.random.default_rng(0)
    m, c, J = 2.0, rng.normal(size=3), rng.normal(size=(3, 3))
    J = J + J.T
    M4 = m4_from_natural_parameters(m, c, J)
    # natural_inertial_parameters.py:153-177, entry by entry
    assert M4[0, 0] == J[0, 0] and M4[0, 1] == m * c[0] + J[0, 1] and M4[0, 2] == -J[0, 1] and M4[0, 3] == J[0, 2]
    assert M4[1, 1] == m + 2 * m * c[1] + J[1, 1] and M4[1, 2] == -(m * c[1] + J[1, 1]) and M4[1, 3] == m * c[2] + J[1, 2]
    assert M4[2, 2] == J[1, 1] and M4[2, 3] == -J[1, 2] and M4[3, 3] == J[2, 2]
    np.testing.assert_array_equal(M4, M4.T)


@pytest.mark.parametrize(
    "name,gen",
    [
        ("lower_limb", lambda t: states.tree_states(t, 50, seed=3)),
        ("full_body", lambda t: states.tree_states(t, 20, seed=4)),
        ("n_link_4", lambda t: states.planar_chain_states(t, 30, seed=5)),
        ("pendulum_force_batch", lambda t: states.hinge_pendulum_states(40, seed=6)),
    ],
)
def test_workload_generators_are_constraint_consistent(name, gen):
    t = ModelTable.load(model_path(name))
    om = OracleModel(model_path(name))
    Q, Qd = gen(t)
    for s in range(0, Q.shape[0], 7):
        assert np.abs(om.holonomic_constraints(Q[s])).max() < 1e-14
        assert np.abs(om.holonomic_constraints_jacobian(Q[s]) @ Qd[s]).max() < 1e-13


def test_knee_generator_keeps_the_closed_loops_closed():
    case = load_case("knee_feikes")
    om = OracleModel(model_path("knee_feikes"))
    Q, Qd = states.rigidly_moved_states(case["Q"][0], 20, seed=9, angle_max=0.3, omega_sigma=0.5)
    for s in range(20):
        assert np.abs(om.holonomic_constraints(Q[s])).max() < 1e-14
        assert np.abs(om.holonomic_constraints_jacobian(Q[s]) @ Qd[s]).max() < 1e-13


def test_flop_counts_match_the_survey():
    expect = {"pendulum_force_batch": (3870, 15673), "knee_feikes": (11787, 47531), "lower_limb": (26538, 106918), "n_link_4": (52481, 210691)}
    for name, (fe, fs) in expect.items():
        t = ModelTable.load(model_path(name))
        assert round(roofline.flops_per_eval(t)) == fe
        assert round(roofline.flops_per_step(t)) == fs
    assert roofline.bytes_per_step(ModelTable.load(model_path("lower_limb"))) == 1536


def test_generic_rk4_mirror_matches_reference_known_answer():
    # reference tests/test_ode_solver.py:6-53
    from bionc_b200.utils import RK4

    expected = [1.0, 1.12392674, 1.27547487, 1.45789044, 1.67480095, 1.9302602, 2.22879841, 2.57547816, 2.97595701, 3.43655736]
    t = np.linspace(0, 1, 10)
    y = RK4(t, lambda t, y: y + t, np.array([1.0]))
    assert np.allclose(y[0], expected, rtol=1e-5)
    y = RK4(t, lambda t, y: y + t, np.array([1.0, 0.0]), normalize_idx=((1,),))
    assert np.allclose(y, np.array([expected, [0.0] + [1.0] * 9]), rtol=1e-5)
    # batch axis: two identical systems evolve identically
    yb = RK4(t, lambda t, y: y + t, np.array([[1.0], [1.0]]))
    assert yb.shape == (2, 1, 10) and np.allclose(yb[0, 0], expected, rtol=1e-5) and np.array_equal(yb[0], yb[1])


def test_natural_types_mirror_the_reference_layout():
    from bionc_b200.bionc_cuda import NaturalCoordinates, NaturalVelocities, SegmentNaturalCoordinates, SegmentNaturalVelocities

    # reference natural_coordinates.py:51-69: slots u, rp, rd, w and v = rp - rd
    Q0 = SegmentNaturalCoordinates.from_components(u=[1, 2, 3.05], rp=[1.1, 1, 3.1], rd=[1.2, 2, 4.1], w=[1.3, 2, 5.1])
    np.testing.assert_array_equal(Q0.u, [1, 2, 3.05])
    np.testing.assert_allclose(Q0.v, [-0.1, -1.0, -1.0])
    Q1 = SegmentNaturalCoordinates.from_components([1.4, 2.1, 3.2], [1.5, 1.1, 3.2], [1.6, 2.2, 4.2], [1.7, 2, 5.3])
    Q = NaturalCoordinates.from_qi((Q0, Q1))
    assert Q.shape == (24, 1) and Q.nb_qi() == 2  # one system is a column, like the reference
    np.testing.assert_array_equal(Q.vector(1).w, [1.7, 2, 5.3])
    np.testing.assert_array_equal(Q.to_array(), np.concatenate([np.asarray(Q0), np.asarray(Q1)]))
    Qd = NaturalVelocities.from_qdoti((SegmentNaturalVelocities.from_components(udot=[0, 0, 0], rpdot=[1, 0, 0], rddot=[0, 1, 0], wdot=[0, 0, 0]),))
    np.testing.assert_array_equal(Qd.vector(0).v, [1, -1, 0])
    Qb = NaturalCoordinates(np.arange(48.0).reshape(2, 24))  # batch of two systems
    assert Qb.nb_qi() == 2 and Qb.vector(1).rp.shape == (2, 3)
    with pytest.raises(ValueError):
        NaturalCoordinates(np.zeros(13))


def test_external_force_set_lowering_matches_the_oracle_kinds():
    fs = ExternalForceSet.empty_from_nb_segment(2)
    fs.add_in_global_local_point(0, np.arange(6.0), point_in_local=[0.1, 0.2, 0.3])
    fs.add_in_global(1, np.arange(6.0) + 1, point_in_global=[1, 2, 3])
    fs.add_in_local(1, np.arange(6.0) + 2, point_in_local=[0, 0.5, 0], transformation_matrix=np.diag([1.0, 2.0, 4.0]))
    fs.add_in_global_on_proximal(0, np.arange(6.0) + 3)
    seg, kind, par = fs.to_arrays()
    assert list(seg) == [0, 0, 1, 1] and list(kind) == [1, 0, 2, 3]
    np.testing.assert_array_equal(par[0, :9], [0, 1, 2, 3, 4, 5, 0.1, 0.2, 0.3])
    np.testing.assert_allclose(par[3, 9:18].reshape(3, 3), np.diag([1.0, 0.5, 0.25]))
    with pytest.raises(ValueError):
        ExternalForceSet(None)
    with pytest.raises(ValueError):
        fs.add_in_local(0, np.zeros(6))


REFERENCE = os.environ.get("BIONC_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "bionc")), reason="the reference checkout is only present in the build container")
def test_model_compiler_reproduces_committed_tables_from_the_reference():
    """ModelTable.from_numpy on freshly loaded reference models == the committed fixtures."""
    import subprocess
    import sys

    code = r'''
import sys, warnings
warnings.filterwarnings("ignore")
sys.path[:0] = [%r, %r, %r]
import numpy as np
from bionc.bionc_numpy import BiomechanicalModel
from bionc_b200.bionc_cuda.model_table import ModelTable
for fn, name in (("pendulum.nmod", "pendulum_root"), ("pendulum_with_force.nmod", "pendulum_force_batch")):
    t = ModelTable.from_numpy(BiomechanicalModel.load(%r + "/" + fn))
    ref = ModelTable.load(%r + "/models/" + name + ".npz")
    for k in ModelTable._ARRAYS:
        assert np.array_equal(getattr(t, k), getattr(ref, k)), (name, k)
# a model with a ground segment (protocols/biomechanical_model.py:283-311): indices shift, the ground is skipped
import importlib.util, os
spec = importlib.util.spec_from_file_location("nlp", %r + "/examples/forward_dynamics/n_link_pendulum.py")
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
from bionc.bionc_numpy import NaturalCoordinates, NaturalVelocities
sys.path.insert(0, %r)
from bionc_oracle import OracleModel
gm = mod.build_n_link_pendulum(nb_segments=4)
gm.remove_joint(name="hinge_0"); gm.remove_joint(name="hinge_1"); gm.set_ground_segment(name="pendulum_0")
gt = ModelTable.from_numpy(gm)
assert gt.nb_segments == 3 and gt.segment_names == ["pendulum_1", "pendulum_2", "pendulum_3"], gt.segment_names
assert gt.nb_holonomic_constraints == gm.nb_holonomic_constraints == 28
Qg = np.linspace(0.1, 0.9, 36); Qdg = np.linspace(0, 0.02, 36)
ref_K = gm.holonomic_constraints_jacobian(NaturalCoordinates(Qg))
om = OracleModel(gt)
assert np.allclose(om.holonomic_constraints_jacobian(Qg), ref_K, rtol=0, atol=1e-14)
print("ground ok")
# a reference ExternalForceSet lowers to the same arrays as the mirror's own construction API
from bionc.bionc_numpy import ExternalForceSet as RefSet
from bionc_b200.bionc_cuda import ExternalForceSet as CudaSet
rs = RefSet.empty_from_nb_segment(2)
rs.add_in_global_local_point(segment_index=0, external_force=np.arange(6.0), point_in_local=np.array([0.1, 0.2, 0.3]))
rs.add_in_global(segment_index=1, external_force=np.arange(6.0) + 1, point_in_global=np.array([1.0, 2.0, 3.0]))
rs.add_in_local(segment_index=1, external_force=np.arange(6.0) + 2, point_in_local=np.array([0, 0.5, 0]), transformation_matrix=np.diag([1.0, 2.0, 4.0]))
cs = CudaSet.empty_from_nb_segment(2)
cs.add_in_global_local_point(0, np.arange(6.0), point_in_local=[0.1, 0.2, 0.3])
cs.add_in_global(1, np.arange(6.0) + 1, point_in_global=[1, 2, 3])
cs.add_in_local(1, np.arange(6.0) + 2, point_in_local=[0, 0.5, 0], transformation_matrix=np.diag([1.0, 2.0, 4.0]))
for a, b in zip(CudaSet.from_reference(rs).to_arrays(), cs.to_arrays()):
    assert np.array_equal(a, b)
print("fext ok")
# error conventions at compile time: a model without inertia (lower_limb.nc as shipped) raises ValueError
try:
    ModelTable.from_numpy(BiomechanicalModel.load(%r + "/examples/models/lower_limb.nc"))
except ValueError as e:
    print("ValueError ok")
print("tables ok")
''' % (os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "stubs"), REFERENCE, os.path.join(GOLDEN, "..", ".."), REFERENCE, GOLDEN, REFERENCE, os.path.join(GOLDEN, "..", "..", "oracle"), REFERENCE)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(os.path.dirname(GOLDEN)))
    for token in ("tables ok", "ValueError ok", "ground ok", "fext ok"):
        assert token in res.stdout, res.stdout + res.stderr


def test_model_table_format_3_keeps_the_joint_list(tmp_path):
    """Format 3 adds every joint's (parent, child) in the reference's joint order -- what joint generalized forces are
    indexed by; older files load with an empty list."""
    t = ModelTable.load(model_path("lower_limb"))
    assert t.all_joint_names == [] and t.all_joint_parent.size == 0  # committed before format 3
    t.all_joint_names = ["free_joint_PELVIS", "hip", "knee", "ankle"]
    t.all_joint_parent, t.all_joint_child = np.array([-1, 0, 1, 2]), np.array([0, 1, 2, 3])
    f = tmp_path / "ll3.npz"
    t.validate().save(str(f))
    t2 = ModelTable.load(str(f))
    assert t2.all_joint_names == t.all_joint_names
    assert np.array_equal(t2.all_joint_parent, [-1, 0, 1, 2]) and np.array_equal(t2.all_joint_child, [0, 1, 2, 3])
    assert t2.all_joint_parent.dtype == np.int32


def test_model_compiler_records_the_reference_joint_order():
    """from_numpy on the live reference (when it is installed in baseline/_ref): the joint list of the lower limb."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "baseline"))
    import reference_models as rm

    if not rm.available():
        pytest.skip("baseline/_ref not installed")
    import warnings

    warnings.filterwarnings("ignore")
    t = ModelTable.from_numpy(rm.lower_limb_model())
    assert t.all_joint_names == ["free_joint_PELVIS", "hip", "knee", "ankle"]
    assert t.all_joint_parent.tolist() == [-1, 0, 1, 2] and t.all_joint_child.tolist() == [0, 1, 2, 3]
    committed = ModelTable.load(model_path("lower_limb"))
    for k in ModelTable._ARRAYS:
        assert np.array_equal(getattr(t, k), getattr(committed, k)), k
