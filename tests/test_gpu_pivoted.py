"""The dense pivoted-LU kernel (bionc_cuda_forward_dynamics_dense): the reference's own algorithm
(augmented matrix + LU with partial pivoting, biomechanical_model.py:335-349, :423-426) on the device.
It is the fallback for systems the fast block-elimination kernel flags, and an algorithmically independent
second GPU path: both must agree with the reference goldens and with the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import golden_case_names  # noqa: E402
from helpers import fext_list, load_case, model_path, rel_err  # noqa: E402
from test_gpu_parity import SINGULAR_KKT, _fext, _model  # noqa: E402
from test_gpu_random_models import random_states, random_table  # noqa: E402

CASES = golden_case_names()
TOL_EVAL = 1e-10
PIVOTED = 8


@pytest.mark.parametrize("name", CASES)
def test_dense_kernel_matches_reference(name):
    if name in SINGULAR_KKT:
        pytest.skip("reference KKT matrix is singular for this model")
    case = load_case(name)
    model = _model(name)
    fext = _fext(model, case)
    qdd, lam, status = model.forward_dynamics(case["Q"], case["Qdot"], external_forces=fext, return_status=True, pivoting="always")
    assert (status == PIVOTED).all()
    for s in range(case["Q"].shape[0]):
        assert rel_err(qdd[s], case["ref_Qddot"][s]) < TOL_EVAL, f"Qddot system {s}"
        assert rel_err(lam[s], case["ref_lam"][s]) < TOL_EVAL, f"lambda system {s}"
    a, b = case["stab_alpha_beta"]
    qdd, lam = model.forward_dynamics(case["Q"], case["Qdot"], external_forces=fext, stabilization=dict(alpha=a, beta=b), pivoting="always")
    for s in range(case["Q"].shape[0]):
        assert rel_err(qdd[s], case["ref_Qddot_stab"][s]) < TOL_EVAL
        assert rel_err(lam[s], case["ref_lam_stab"][s]) < TOL_EVAL


def test_dense_kernel_with_per_system_inertial_parameters():
    case = load_case("full_body_perturbed")
    model = _model("full_body")
    J = case["seg_J"]
    ip = np.concatenate(
        [case["seg_mass"][..., None], case["seg_com"],
         np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1)
    qdd, lam = model.forward_dynamics(case["Q"], case["Qdot"], inertial_parameters=ip, pivoting="always")
    for s in range(case["Q"].shape[0]):
        assert rel_err(qdd[s], case["ref_Qddot"][s]) < TOL_EVAL
        assert rel_err(lam[s], case["ref_lam"][s]) < TOL_EVAL


def test_fast_and_dense_paths_agree_on_a_large_batch():
    """Two algorithmically independent GPU paths on 20 000 lower-limb systems (more systems than workspace slots,
    so the CTAs of the dense kernel loop)."""
    from bionc_b200.utils import states as st

    model = _model("lower_limb")
    Q, Qd = st.tree_states(model.table, 20_000, seed=5)
    a = model.forward_dynamics(Q, Qd, pivoting="never")
    b = model.forward_dynamics(Q, Qd, pivoting="always")
    assert rel_err(a[0], b[0]) < 1e-9 and rel_err(a[1], b[1]) < 1e-9


def test_auto_pivoting_on_indefinite_pseudo_inertia():
    """Random symmetric (indefinite) J: where a 3x3 inverse of the fast path (inertia tensor about the centre of mass,
    joint-level pivots) meets a small determinant, "auto" re-solves with the dense kernel.  Every system whose KKT matrix is reasonably conditioned must then match
    the oracle's pivoted LU -- no escape through the status flag."""
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda import BiomechanicalModel

    rng = np.random.default_rng(0)
    checked = pivoted = 0
    for trial in range(40):
        N = int(rng.integers(1, 7))
        table = random_table(5000 + trial, N, int(rng.integers(0, 3)) if N > 1 else 0)
        for i in range(N):
            A = rng.normal(size=(3, 3)) * 0.3
            table.seg_J[i] = A + A.T
        table.validate()
        model = BiomechanicalModel.from_table(table)
        om = OracleModel(table)
        Q, Qd = random_states(table, 3, trial)
        qdd, lam, status = model.forward_dynamics(Q, Qd, return_status=True)
        assert not (status & 4).any()
        for s in range(3):
            cond = np.linalg.cond(om.augmented_mass_matrix(Q[s]))
            if cond > 1e10:
                continue
            rq, rl = om.forward_dynamics(Q[s], Qd[s])
            # fast path: 1e-10 up to cond 1e4; beyond, the unpivoted block elimination of an INDEFINITE joint-level
            # matrix loses digits with the condition number (eps * cond * growth; trial 17: cond(S') = 9e4, 2e-10).
            # A system that went through the dense kernel shares the oracle's algorithm and is held to the same bound
            tol = 1e-10 * max(1.0, cond / 1e4)
            assert rel_err(qdd[s], rq) < tol and rel_err(lam[s], rl) < tol, (trial, N, s, cond, int(status[s]))
            checked += 1
            pivoted += int(status[s] == PIVOTED)
    assert checked > 100


def test_exactly_singular_matrix_raises_like_numpy():
    """The same spherical joint twice gives two identical constraint blocks: the KKT matrix is exactly singular.
    numpy.linalg.solve either raises LinAlgError or returns rounding noise for it (the duplicated rows cancel to
    0 or to ~1e-17 during the elimination); the kernel must report the system either way and never return a
    clean status."""
    from bionc_b200.bionc_cuda import BiomechanicalModel
    from bionc_b200.bionc_cuda.model_table import BLK_LIN3
    from test_gpu_random_models import assemble_table

    base = random_table(77, 2, 0)
    sph = (BLK_LIN3, np.concatenate([[0, 0, 1.0, 0], [0, 1.0, 0, 0], np.zeros(3)]))
    table = assemble_table(base.seg_phi_const, base.seg_geometry, base.seg_mass, base.seg_com, base.seg_J,
                           [(0, 1, [sph]), (0, 1, [sph])])
    model = BiomechanicalModel.from_table(table)
    Q, Qd = random_states(table, 4, 3)
    _, _, st_fast = model.forward_dynamics(Q, Qd, return_status=True, pivoting="never")
    assert (st_fast & 5).all()  # zero or tiny pivot in the unpivoted LDL^T
    qdd, lam, st = model.forward_dynamics(Q, Qd, return_status=True)
    assert (st & PIVOTED).all()
    for s in range(4):
        if st[s] & 1:
            with pytest.raises(np.linalg.LinAlgError):
                model.forward_dynamics(Q[s], Qd[s])
        else:  # rounding left a ~1e-17 pivot, as it does in LAPACK: huge multipliers, like the reference returns
            assert np.abs(lam[s]).max() > 1e6 or not np.isfinite(lam[s]).all()


def test_stagewise_pivoted_rk4_agrees_with_the_fused_kernel():
    """The fallback integrator (four dense pivoted-LU launches per step) against the fused kernel and the golden
    trajectory on a model that needs no pivoting: same method, two independent solvers."""
    case = load_case("lower_limb")
    model = _model("lower_limb")
    t = case["traj_cfg_t"]
    nt = case["traj_cfg_final"].shape[0]
    Q = torch.from_numpy(case["Q"][:nt]).cuda()
    Qd = torch.from_numpy(case["Qdot"][:nt]).cuda()
    Qf, Qdf, traj_f, lam_f = model.rk4(Q, Qd, t=t, normalize=True, save_every=10, return_lambdas=True)
    Qs, Qds = Q.clone(), Qd.clone()
    traj_s = torch.zeros_like(traj_f)
    lam_s = torch.zeros_like(lam_f)
    status = torch.zeros((nt,), dtype=torch.int32, device="cuda")
    sel = torch.arange(nt, device="cuda")
    model._rk4_stagewise_pivoted(sel, (Q, Qd), t[1:] - t[:-1], True, None, None, None, Qs, Qds, traj_s, 10, lam_s, status)
    assert (status == PIVOTED).all()
    assert rel_err(Qs.cpu().numpy(), Qf.cpu().numpy()) < 1e-9 and rel_err(Qds.cpu().numpy(), Qdf.cpu().numpy()) < 1e-9
    assert rel_err(traj_s.cpu().numpy(), traj_f.cpu().numpy()) < 1e-9
    assert rel_err(lam_s.cpu().numpy(), lam_f.cpu().numpy()) < 1e-8
    for s in range(nt):
        assert rel_err(np.concatenate([Qs[s].cpu().numpy(), Qds[s].cpu().numpy()]), case["traj_cfg_final"][s]) < 1e-8


def test_rk4_auto_pivoting_on_flagged_systems():
    """Indefinite models whose fused integration raises a flag are re-integrated with the pivoted kernel and must
    then follow the oracle's (LU) trajectory; unflagged systems keep the bits of the fused result."""
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda import BiomechanicalModel

    rng = np.random.default_rng(1)
    t = np.linspace(0.0, 0.01, 11)
    compared = 0
    for trial in range(60):
        N = int(rng.integers(1, 5))
        table = random_table(7000 + trial, N, int(rng.integers(0, 2)) if N > 1 else 0)
        for i in range(N):
            A = rng.normal(size=(3, 3)) * 0.3
            table.seg_J[i] = A + A.T
        table.validate()
        model = BiomechanicalModel.from_table(table)
        Q, Qd = random_states(table, 4, trial)
        Qn, Qdn, st_n = model.rk4(Q, Qd, t=t, normalize=True, return_status=True)
        Qa, Qda, st_a = model.rk4(Q, Qd, t=t, normalize=True, return_status=True, pivoting="auto")
        clean = st_n == 0
        assert np.array_equal(Qa[clean], Qn[clean]) and np.array_equal(Qda[clean], Qdn[clean])
        assert ((st_a[~clean] & PIVOTED) != 0).all() and (st_a[clean] == 0).all()
        if (~clean).any() and compared < 6:
            om = OracleModel(table)
            for s in np.nonzero(~clean)[0]:
                if st_a[s] != PIVOTED:
                    continue  # singular or non-finite along the way: the oracle fails there as well
                try:
                    rq, rqd = om.rk4(Q[s], Qd[s], t, normalize=True)
                except np.linalg.LinAlgError:
                    continue
                ref = np.concatenate([rq, rqd])
                if not np.isfinite(ref).all() or np.abs(ref).max() > 1e6:
                    continue  # unstable trajectory: nothing to compare
                assert rel_err(np.concatenate([Qa[s], Qda[s]]), ref) < 1e-6, (trial, s)
                compared += 1
    print("re-integrated systems compared with the oracle:", compared)


def _zero_pivot_table(N):
    """Segment 0 is a thin rod along v (all its second moments on the v axis: J = diag(0, 1, 0), c on the axis), held
    by a ground hinge about x.  Its inertia tensor about the centre of mass, Ic = 0.25 (|v|^2 I - v v^T), is singular
    (no inertia about the rod's own axis), so the segment-level 3x3 inverse of the fast path has a zero determinant,
    while the KKT matrix is well conditioned (cond 20 .. 700): the hinge takes the spin about the rod away.  G itself
    is singular here (zero rows for u and w).  numpy's pivoted LU has no problem with it."""
    from bionc_b200.bionc_cuda.model_table import BLK_DOT, BLK_LIN3
    from test_gpu_random_models import assemble_table

    base = random_table(1, N, 0)
    mass, com, J = base.seg_mass.copy(), base.seg_com.copy(), base.seg_J.copy()
    mass[0], com[0], J[0] = 3.0, np.array([0.0, -0.5, 0.0]), np.diag([0.0, 1.0, 0.0])
    ground_hinge = [(BLK_LIN3, np.concatenate([np.zeros(4), [0, 1.0, 0, 0], np.zeros(3)])),
                    (BLK_DOT, np.concatenate([[1.0, 0, 0, 0], [0, 1.0, -1.0, 0], [0.0]])),
                    (BLK_DOT, np.concatenate([[1.0, 0, 0, 0], [0, 0, 0, 1.0], [0.0]]))]
    sph = (BLK_LIN3, np.concatenate([[0, 0, 1.0, 0], [0, 1.0, 0, 0], np.zeros(3)]))
    joints = [(-1, 0, ground_hinge)] + [(i - 1, i, [sph]) for i in range(1, N)]
    return assemble_table(base.seg_phi_const, base.seg_geometry, mass, com, J, joints)


@pytest.mark.parametrize("N", [1, 3])
def test_zero_pivot_of_the_fast_path_is_resolved_by_the_dense_kernel(N):
    from bionc_oracle import OracleModel

    from bionc_b200.bionc_cuda import BiomechanicalModel

    table = _zero_pivot_table(N)
    model = BiomechanicalModel.from_table(table)
    om = OracleModel(table)
    Q, Qd = random_states(table, 5, 11)
    _, _, st_fast = model.forward_dynamics(Q, Qd, return_status=True, pivoting="never")
    assert (st_fast != 0).all()
    qdd, lam, st = model.forward_dynamics(Q, Qd, return_status=True)
    assert (st == PIVOTED).all()
    for s in range(5):
        assert np.linalg.cond(om.augmented_mass_matrix(Q[s])) < 1e6
        rq, rl = om.forward_dynamics(Q[s], Qd[s])
        assert rel_err(qdd[s], rq) < TOL_EVAL and rel_err(lam[s], rl) < TOL_EVAL
    # single system: no LinAlgError, reference shapes
    a, l = model.forward_dynamics(Q[0], Qd[0])
    assert a.shape == (12 * N, 1) and rel_err(a[:, 0], om.forward_dynamics(Q[0], Qd[0])[0]) < TOL_EVAL
    # fused RK4 flags the systems, "auto" re-integrates them along the oracle's trajectory
    t = np.linspace(0.0, 0.02, 21)
    _, _, st_n = model.rk4(Q, Qd, t=t, normalize=True, return_status=True)
    assert (st_n != 0).all()
    Qa, Qda, traj, lam_a, st_a = model.rk4(Q, Qd, t=t, normalize=True, save_every=10, return_lambdas=True, return_status=True, pivoting="auto")
    assert (st_a == PIVOTED).all() and traj.shape == (2, 5, 24 * N)
    for s in range(5):
        rq, rqd = om.rk4(Q[s], Qd[s], t, normalize=True)
        assert rel_err(np.concatenate([Qa[s], Qda[s]]), np.concatenate([rq, rqd])) < 1e-8
        mq, mqd = om.rk4(Q[s], Qd[s], t[:11], normalize=True)
        assert rel_err(traj[0, s], np.concatenate([mq, mqd])) < 1e-8
    assert np.array_equal(traj[1], np.concatenate([Qa, Qda], axis=1))
