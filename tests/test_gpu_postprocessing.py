"""Rows either side of the hot path against outputs of the unmodified reference
(tests/golden/cases): the single-segment DAE (natural_segment.py:574-633), marker / centre-of-mass kinematics
(biomechanical_model_markers.py:32-79), energies (biomechanical_model.py:98-152) and the per-frame residual
norms of bionc_numpy/time_series_utils.py, all evaluated by CUDA kernels through bionc_cuda_eval."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import golden_case_names  # noqa: E402
from helpers import load_case, model_path, rel_err  # noqa: E402
from test_gpu_parity import _model  # noqa: E402

CASES = golden_case_names()
TOL = dict(rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_kinematics_energies_and_residual_norms_match_reference(name):
    case = load_case(name)
    model = _model(name)
    Q, Qd = case["Q"], case["Qdot"]
    B = Q.shape[0]
    assert model.nb_markers == case["ref_markers"].shape[2]
    np.testing.assert_allclose(model.markers(Q), case["ref_markers"], **TOL)
    np.testing.assert_allclose(model.center_of_mass_position(Q), case["ref_com"], **TOL)
    np.testing.assert_allclose(model.kinetic_energy(Qd), case["ref_kinetic"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(model.potential_energy(Q), case["ref_potential"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(model.lagrangian(Q, Qd), case["ref_kinetic"] - case["ref_potential"], rtol=1e-12, atol=1e-12)
    norms = model.constraint_residual_norms(Q)
    np.testing.assert_allclose(norms[:, 0], case["ref_total_rigid"], **TOL)
    np.testing.assert_allclose(norms[:, 1], case["ref_total_joint"], **TOL)
    # one system: the reference's shapes
    mk = model.markers(Q[0])
    assert mk.shape == (3, model.nb_markers, 1)
    np.testing.assert_allclose(mk[:, :, 0], case["ref_markers"][0], **TOL)
    com = model.center_of_mass_position(Q[0].reshape(-1, 1))
    assert com.shape == (3, model.nb_segments, 1)
    assert isinstance(model.kinetic_energy(Qd[0]), float) and isinstance(model.potential_energy(Q[0]), float)
    assert model.constraint_residual_norms(Q[0]).shape == (2,)
    # torch in, torch out (stays on the device)
    out = model.markers(torch.from_numpy(Q).cuda())
    assert out.is_cuda and tuple(out.shape) == (B, 3, model.nb_markers)


@pytest.mark.parametrize("name", CASES)
def test_single_segment_dae_matches_reference(name):
    """NaturalSegment.differential_algebraic_equation, every segment of every golden model."""
    case = load_case(name)
    model = _model(name)
    Q, Qd = case["Q"], case["Qdot"]
    G = model.mass_matrix
    for i in range(model.nb_segments):
        seg = model.segment(i)
        assert seg.name == model.segment_names[i]
        Qi, Qdi = Q[:, 12 * i : 12 * i + 12], Qd[:, 12 * i : 12 * i + 12]
        qdd, lam, status = seg.differential_algebraic_equation(Qi, Qdi, return_status=True)
        assert qdd.shape == (Q.shape[0], 12) and lam.shape == (Q.shape[0], 6)
        assert not (status & 3).any()
        for s in range(Q.shape[0]):
            assert rel_err(qdd[s], case["ref_dae_Qddot"][s, i]) < 1e-10, (i, s)
            assert rel_err(lam[s], case["ref_dae_lam"][s, i]) < 1e-10, (i, s)
        a, l = seg.differential_algebraic_equation(Qi[0], Qdi[0])
        assert a.shape == (12,) and l.shape == (6,)  # the reference returns (12,), (6,)
        np.testing.assert_array_equal(a, qdd[0])
        np.testing.assert_allclose(seg.mass_matrix, G[12 * i : 12 * i + 12, 12 * i : 12 * i + 12], rtol=0, atol=0)
        np.testing.assert_allclose(seg.gravity_force(), case["ref_gravity"][12 * i : 12 * i + 12], rtol=0, atol=1e-13)
        np.testing.assert_allclose(seg.rigid_body_constraint(Qi), case["ref_phi_r"][:, 6 * i : 6 * i + 6], **TOL)
        np.testing.assert_allclose(seg.center_of_mass_position(Qi), case["ref_com"][:, :, i], **TOL)


def test_single_segment_dae_with_stabilization_matches_the_oracle():
    """The reference's own Baumgarte branch of the segment DAE raises (natural_segment.py:617); the intended
    formula is pinned against the oracle's restatement only."""
    from bionc_oracle import OracleModel

    case = load_case("lower_limb")
    model = _model("lower_limb")
    om = OracleModel(model_path("lower_limb"))
    stab = dict(alpha=0.7, beta=1.3)
    rng = np.random.default_rng(3)
    for i in range(model.nb_segments):
        Qi = case["Q"][:, 12 * i : 12 * i + 12] + 1e-3 * rng.normal(size=(case["Q"].shape[0], 12))  # off the constraint manifold
        Qdi = case["Qdot"][:, 12 * i : 12 * i + 12]
        qdd, lam = model.segment(i).differential_algebraic_equation(Qi, Qdi, stabilization=stab)
        for s in range(Qi.shape[0]):
            ra, rl = om.segment_differential_algebraic_equation(i, Qi[s], Qdi[s], stabilization=stab)
            assert rel_err(qdd[s], ra) < 1e-10 and rel_err(lam[s], rl) < 1e-10


def test_time_series_utils_on_a_device_trajectory():
    """time_series_utils.total_rigid_body_constraints / total_joint_constraints over the snapshots of the fused
    integrator, without leaving the device, against the oracle frame by frame; reference layout (n, frames) too."""
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda import time_series_utils as tsu

    case = load_case("lower_limb")
    model = _model("lower_limb")
    om = OracleModel(model_path("lower_limb"))
    n = model.nb_Q
    Q = torch.from_numpy(case["Q"]).cuda()
    Qd = torch.from_numpy(case["Qdot"]).cuda()
    _, _, traj = model.rk4(Q, Qd, h=0.01, n_steps=50, normalize=True, save_every=10)  # (5, B, 2n) on the device
    tr = tsu.total_rigid_body_constraints(model, traj)
    tj = tsu.total_joint_constraints(model, traj)
    assert tr.is_cuda and tuple(tr.shape) == (5, Q.shape[0]) and tuple(tj.shape) == (5, Q.shape[0])
    host = traj.cpu().numpy()
    for f in range(5):
        for s in range(Q.shape[0]):
            assert abs(float(tr[f, s]) - om.total_rigid_body_constraints(host[f, s, :n])) < 1e-13
            assert abs(float(tj[f, s]) - om.total_joint_constraints(host[f, s, :n])) < 1e-13
    # reference layout: Q (n, frames) of one system -> (frames,), and the residual vectors (6N, frames)
    Qref = host[:, 0, :n].T
    assert Qref.shape == (n, 5)
    np.testing.assert_allclose(tsu.total_rigid_body_constraints(model, Qref), tr[:, 0].cpu().numpy(), rtol=0, atol=1e-15)
    R = tsu.TimeSeriesUtils.rigid_body_constraints(model, Qref)
    assert R.shape == (model.nb_rigid_body_constraints, 5)
    np.testing.assert_allclose(np.sqrt((R**2).sum(axis=0)), tr[:, 0].cpu().numpy(), rtol=1e-10, atol=1e-15)
    assert tsu.joint_constraints(model, Qref).shape == (model.nb_joint_constraints, 5)
    # drift stays small over the run
    assert float(tr.max()) < 1e-6 and float(tj.max()) < 1e-6


def test_postprocessing_at_scale_properties():
    """1e6 systems: rigid transport invariance of the residual norms and of the kinetic energy; potential energy and
    marker heights shift by the translation."""
    from bionc_b200.utils import states as st

    model = _model("lower_limb")
    B = 1_000_000
    Q, Qd = st.tree_states(model.table, B, seed=9)
    Qt, Qdt = torch.from_numpy(Q).cuda(), torch.from_numpy(Qd).cuda()
    norms = model.constraint_residual_norms(Qt)
    assert float(norms.max()) < 1e-12
    shift = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64, device="cuda")
    Q2 = Qt.clone().view(B, model.nb_segments, 4, 3)
    Q2[:, :, 1] += shift  # rp
    Q2[:, :, 2] += shift  # rd
    Q2 = Q2.view(B, -1)
    mass = float(model.table.seg_mass.sum())
    dE = model.potential_energy(Q2) - model.potential_energy(Qt)
    assert float((dE - mass * 0.5).abs().max()) < 1e-10
    mk, mk2 = model.markers(Qt), model.markers(Q2)
    assert float((mk2 - mk - shift.view(1, 3, 1)).abs().max()) < 1e-12
    com = model.center_of_mass_position(Qt)
    E = model.potential_energy(Qt)
    mz = torch.from_numpy(model.table.seg_mass).cuda()
    assert float((E - (com[:, 2, :] * mz).sum(dim=1)).abs().max()) < 1e-10
    T = model.kinetic_energy(Qdt)
    assert float(T.min()) >= 0.0
    G = torch.from_numpy(model.mass_matrix).cuda()
    Tref = 0.5 * ((Qdt[:4096] @ G) * Qdt[:4096]).sum(dim=1)
    assert float((T[:4096] - Tref).abs().max()) < 1e-10 * max(1.0, float(Tref.abs().max()))
