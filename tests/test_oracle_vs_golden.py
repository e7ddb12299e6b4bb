"""Pin the CPU oracle (oracle/bionc_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/cases/*.npz, written by tests/golden/generate_golden.py in the build container) and
against literal known-answer vectors transcribed from the reference's own test-suite."""
import os

import numpy as np
import pytest

from bionc_oracle import OracleModel, rk4_generic
from conftest import golden_case_names
from helpers import GOLDEN, fext_list, load_case, model_path, rel_err

CASES = golden_case_names()


@pytest.mark.parametrize("name", CASES)
def test_oracle_components_match_reference(name):
    case = load_case(name)
    om = OracleModel(model_path(name))
    fext = fext_list(case)
    np.testing.assert_allclose(om.mass_matrix, case["ref_mass_matrix"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(om.gravity_forces(), case["ref_gravity"], rtol=0, atol=1e-13)
    for s in range(case["Q"].shape[0]):
        Q, Qd = case["Q"][s], case["Qdot"][s]
        tol = dict(rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(om.rigid_body_constraints(Q), case["ref_phi_r"][s], **tol)
        np.testing.assert_allclose(om.joint_constraints(Q), case["ref_phi_j"][s], **tol)
        np.testing.assert_allclose(om.joint_constraints_jacobian(Q), case["ref_K_j"][s], **tol)
        np.testing.assert_allclose(om.holonomic_constraints(Q), case["ref_phi_h"][s], **tol)
        np.testing.assert_allclose(om.holonomic_constraints_jacobian(Q), case["ref_K_h"][s], **tol)
        np.testing.assert_allclose(om.holonomic_constraints_acceleration_bias(Qd), case["ref_bias"][s], **tol)
        np.testing.assert_allclose(om.natural_external_forces(Q, fext), case["ref_fext"][s], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(om.kinetic_energy(Qd), case["ref_kinetic"][s], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(om.potential_energy(Q), case["ref_potential"][s], rtol=1e-13, atol=1e-13)
        # post-processing rows and the single-segment DAE (natural_segment.py:574-633)
        np.testing.assert_allclose(om.center_of_mass_position(Q), case["ref_com"][s], **tol)
        np.testing.assert_allclose(om.markers(Q), case["ref_markers"][s], **tol)
        np.testing.assert_allclose(om.total_rigid_body_constraints(Q), case["ref_total_rigid"][s], **tol)
        np.testing.assert_allclose(om.total_joint_constraints(Q), case["ref_total_joint"][s], **tol)
        for i in range(om.nb_segments):
            a, l = om.segment_differential_algebraic_equation(i, Q[12 * i : 12 * i + 12], Qd[12 * i : 12 * i + 12])
            assert rel_err(a, case["ref_dae_Qddot"][s, i]) < 1e-10 and rel_err(l, case["ref_dae_lam"][s, i]) < 1e-10


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_dynamics_matches_reference(name):
    case = load_case(name)
    om = OracleModel(model_path(name))
    fext = fext_list(case)
    a, b = case["stab_alpha_beta"]
    for s in range(case["Q"].shape[0]):
        qdd, lam = om.forward_dynamics(case["Q"][s], case["Qdot"][s], fext)
        assert rel_err(qdd, case["ref_Qddot"][s]) < 1e-10
        assert rel_err(lam, case["ref_lam"][s]) < 1e-10
        qdd, lam = om.forward_dynamics(case["Q"][s], case["Qdot"][s], fext, stabilization=dict(alpha=a, beta=b))
        assert rel_err(qdd, case["ref_Qddot_stab"][s]) < 1e-10
        assert rel_err(lam, case["ref_lam_stab"][s]) < 1e-10


TRAJ = [
    ("pendulum_x_revolute", "example", 1),
    ("pendulum_force_no_equilibrium", "example", 1),
    ("double_pendulum_force_no_equilibrium", "example", 1),
    ("pendulum_force_batch", "cfg", 1),
    ("knee_feikes", "cfg", 1),
    ("lower_limb", "fine", 1),
    ("n_link_4", "short", 1),
    ("free_box", "example", 1),
]


@pytest.mark.parametrize("name,tag,nsys", TRAJ)
def test_oracle_rk4_matches_reference(name, tag, nsys):
    case = load_case(name)
    om = OracleModel(model_path(name))
    fext = fext_list(case)
    t = case[f"traj_{tag}_t"]
    for s in range(nsys):
        Qf, Qdf = om.rk4(case["Q"][s], case["Qdot"][s], t, normalize=bool(case[f"traj_{tag}_normalize"]), fext=fext)
        assert rel_err(np.concatenate([Qf, Qdf]), case[f"traj_{tag}_final"][s]) < 1e-8


# ---------------------------------------------------------------------------------------------
# Literal known answers transcribed from the reference's own tests (file:line in each comment)
# ---------------------------------------------------------------------------------------------
def test_literal_test_joints_revolute():
    # reference tests/test_joints.py:334-377
    case = load_case("tj_revolute")
    om = OracleModel(model_path("tj_revolute"))
    Q = case["Q"][0]
    np.testing.assert_almost_equal(om.joint_constraints(Q), [-0.3, 0.9, 0.9, -5.85, -6.762893], decimal=6)
    K = om.joint_constraints_jacobian(Q)
    np.testing.assert_almost_equal(K[3, 0:12], [-0.1, -1.1, -1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0], decimal=6)
    np.testing.assert_almost_equal(K[4, 0:12], [0, 0, 0, 1.7, 2.0, 5.3, -1.7, -2.0, -5.3, 0, 0, 0], decimal=6)
    np.testing.assert_almost_equal(K[3, 12:24], [0, 0, 0, 1.0, 2.0, 3.05, -1.0, -2.0, -3.05, 0, 0, 0], decimal=6)
    np.testing.assert_almost_equal(K[4, 12:24], [0, 0, 0, 0, 0, 0, 0, 0, 0, -0.1, -1.0, -1.0], decimal=6)
    b = om.holonomic_constraints_acceleration_bias(Q)  # the reference test feeds Q1, Q2 as velocities (:371)
    r0 = om.blk_row_h[0]
    np.testing.assert_almost_equal(b[r0 : r0 + 5], [0, 0, 0, -10.7, -14.94], decimal=6)


def test_literal_test_joints_universal():
    # reference tests/test_joints.py:378-420
    case = load_case("tj_universal")
    om = OracleModel(model_path("tj_universal"))
    Q = case["Q"][0]
    np.testing.assert_almost_equal(om.joint_constraints(Q), [-0.3, 0.9, 0.9, 20.943939], decimal=6)
    b = om.holonomic_constraints_acceleration_bias(Q)
    r0 = om.blk_row_h[0]
    np.testing.assert_almost_equal(b[r0 : r0 + 4], [0, 0, 0, 43.73], decimal=6)


def test_literal_single_pendulum_final_states():
    # reference tests/test_ultimate_examples.py:86-144 (500 RK4 steps, no renormalisation)
    expected = {
        "pendulum_x_revolute": [1, 0, 0, 0, -3.297850e-15, -1.641582e-14, 0, -5.151967e-01, 8.568669e-01, 0, 8.568669e-01, 5.151967e-01],
        "pendulum_y_revolute": [5.151967e-01, 0, -8.568669e-01, -3.661819e-16, 4.391503e-31, -7.966007e-15, -3.427154e-16, -1, -9.281921e-15, 8.568669e-01, 0, 5.151967e-01],
        "pendulum_z_revolute": [0, -8.568669e-01, -5.151967e-01, -2.328333e-17, 1.398126e-15, -8.541665e-15, -9.373376e-18, -5.151967e-01, 8.568669e-01, 1, 0, 0],
    }
    for name, exp in expected.items():
        case = load_case(name)
        om = OracleModel(model_path(name))
        Qf, _ = om.rk4(case["Q"][0], case["Qdot"][0], case["traj_example_t"], normalize=False)
        np.testing.assert_almost_equal(Qf, exp, decimal=6)


def test_literal_pendulum_with_force_no_equilibrium():
    # reference tests/test_ultimate_examples.py:214-243 (600 RK4 steps with renormalisation)
    exp = [1.00000000e00, -1.35070921e-14, -5.74195276e-15, 4.31085832e-15, 6.18406905e-17, 2.11272817e-15,
           6.29819871e-15, -9.42305961e-01, -3.34739950e-01, -7.85747160e-18, -3.34742001e-01, 9.42309818e-01,
           -4.31132552e-30, 7.14524221e-15, -1.75584183e-14, 7.70496015e-16, -5.53946285e-18, 3.70702716e-16,
           2.21909111e-15, 4.28930797e-01, -1.20745434e00, 1.24168228e-17, -1.20746022e00, -4.28933759e-01]
    case = load_case("pendulum_force_no_equilibrium")
    om = OracleModel(model_path("pendulum_force_no_equilibrium"))
    Qf, Qdf = om.rk4(case["Q"][0], case["Qdot"][0], case["traj_example_t"], normalize=True, fext=fext_list(case))
    np.testing.assert_almost_equal(np.concatenate([Qf, Qdf]), exp, decimal=7)


def test_literal_n_link_linspace_state():
    # reference tests/test_forward_dynamics.py:251-334: 4-link pendulum at the non-consistent linspace state
    case = load_case("n_link_4")
    om = OracleModel(model_path("n_link_4"))
    Q, Qd = case["Q_linspace"], case["Qdot_linspace"]
    np.testing.assert_allclose(om.holonomic_constraints_jacobian(Q), case["ref_K_linspace"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(om.holonomic_constraints_acceleration_bias(Qd), case["ref_bias_linspace"], rtol=1e-13, atol=1e-13)
    qdd, _ = om.forward_dynamics(Q, Qd)
    # ill-conditioned there (lambda ~ 1e5, "hard to test it on cross platforms", :492-501): 1e-6 as the reference does
    np.testing.assert_allclose(qdd, case["ref_Qddot_linspace"], rtol=1e-6, atol=1e-6)
    # first values of the literal Qddot block (:283-290)
    np.testing.assert_almost_equal(qdd[:3], case["ref_Qddot_linspace"][:3], decimal=6)


def test_literal_free_box_xdot():
    # reference tests/test_forward_dynamics.py:95-136: single free box at Q = Qdot = linspace(0, 11, 12)
    expected = np.array([-1.97372982e-16, -1.0, -2.0, 1.77635684e-15, -1.59872116e-15, -9.81, -3.0, -3.0, -1.281e01, -9.0, -1.0e01, -1.1e01])
    om = OracleModel(model_path("free_box"))
    qdd, _ = om.forward_dynamics(np.linspace(0, 11, 12), np.linspace(0, 11, 12))
    np.testing.assert_almost_equal(qdd, expected, decimal=6)


def test_literal_ode_solver():
    # reference tests/test_ode_solver.py:6-53: y' = y + t on linspace(0, 1, 10), with and without normalize_idx
    expected = [1.0, 1.12392674, 1.27547487, 1.45789044, 1.67480095, 1.9302602, 2.22879841, 2.57547816, 2.97595701, 3.43655736]
    t = np.linspace(0, 1, 10)
    y = rk4_generic(t, lambda t, y: y + t, np.array([1.0]))
    assert np.allclose(y[0], expected, rtol=1e-5)
    y = rk4_generic(t, lambda t, y: y + t, np.array([1.0, 0.0]), normalize_idx=((1,),))
    assert np.allclose(y, np.array([expected, [0.0] + [1.0] * 9]), rtol=1e-5)


def test_full_body_perturbed_inertial_parameters():
    case = load_case("full_body_perturbed")
    for s in range(case["Q"].shape[0]):
        om = OracleModel(model_path("full_body"), seg_mass=case["seg_mass"][s], seg_com=case["seg_com"][s], seg_J=case["seg_J"][s])
        qdd, lam = om.forward_dynamics(case["Q"][s], case["Qdot"][s])
        assert rel_err(qdd, case["ref_Qddot"][s]) < 1e-10
        assert rel_err(lam, case["ref_lam"][s]) < 1e-10


def test_joint_generalized_forces_against_the_reference_building_blocks():
    """the oracle's joint-reaction force kind and the lowering of a JointGeneralizedForcesList against
    vectors composed from the reference's own methods (tests/golden/generate_joint_forces.py)."""
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda.generalized_force import JointGeneralizedForces, JointGeneralizedForcesList

    with np.load(os.path.join(GOLDEN, "cases_extra", "joint_forces_lower_limb.npz")) as z:
        g = {k: np.array(z[k]) for k in z.files}
    om = OracleModel(model_path("lower_limb"))
    jl = JointGeneralizedForcesList.empty_from_nb_joint(g["all_joint_parent"].size)
    for j, v in zip(g["joint_index"], g["joint_force"]):
        jl.add_generalized_force(int(j), JointGeneralizedForces(v))
    seg, kind, par = jl.to_arrays(g["all_joint_parent"], g["all_joint_child"])
    assert set(kind.tolist()) == {1, 4}  # forces on the children, reactions on the parents
    fext = [(int(a), int(b), c) for a, b, c in zip(seg, kind, par)]
    for s in range(g["Q"].shape[0]):
        assert np.abs(om.natural_external_forces(g["Q"][s], fext) - g["ref_natural_joint_forces"][s]).max() < 1e-12
        qdd, lam = om.forward_dynamics(g["Q"][s], g["Qdot"][s], fext)
        assert rel_err(qdd, g["ref_Qddot"][s]) < 1e-10 and rel_err(lam, g["ref_lam"][s]) < 1e-10
    with pytest.raises(ValueError):
        JointGeneralizedForcesList.empty_from_nb_joint(2).to_arrays(g["all_joint_parent"], g["all_joint_child"])
    with pytest.raises(NotImplementedError):
        JointGeneralizedForces(np.zeros(6), application_point_in_local=[0, 0.1, 0])
