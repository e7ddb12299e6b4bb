"""The drop-in path against the LIVE reference, in one process on the GPU box:

    bionc.bionc_numpy.BiomechanicalModel.load("pendulum.nmod")      (protocols/biomechanical_model.py:263-281)
        -> bionc_cuda.BiomechanicalModel.from_numpy(model)           (the analogue of to_mx(), bionc_numpy/biomechanical_model.py:43-75)
        -> forward_dynamics / rk4 on the device
    beside the unmodified reference's model.forward_dynamics (bionc_numpy/biomechanical_model.py:351-429) and
    bionc.RK4 (utils/ode_solver.py:8-53) on the same seeded states.

The reference is the installed copy in baseline/_ref (pure Python, placed there by __graft_entry__.build(); it travels
to the GPU box with the snapshot).  256 states per configuration for one evaluation, 8 states for 100 RK4 steps.
Every case also records the UNFLOORED errors max|d| / max|ref| (helpers.rel_err floors the denominator at 1) in
gpurun_out/live_reference_errors.json; a copy of a full run is committed as profiles/r02_live_reference_errors.json."""
import json
import os
import sys
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import REPO, rel_err  # noqa: E402

sys.path.insert(0, os.path.join(REPO, "baseline"))
import reference_models as rm  # noqa: E402

TOL_EVAL, TOL_TRAJ = 1e-10, 1e-8
N_EVAL, N_TRAJ, TRAJ_STEPS = 256, 8, 100
ERRORS = {}


def unfloored(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def hinge_swings(model, Q0, count, seed):
    """Consistent states of a one-degree-of-freedom model: its free motion at Q0 is the null space of K(Q0); every
    3-vector of Q0 is rotated about that axis through the origin (the pinned point) by a random angle, with a random rate."""
    from bionc.bionc_numpy import NaturalCoordinates
    from bionc_b200.utils import states as st

    K = np.asarray(model.holonomic_constraints_jacobian(NaturalCoordinates(Q0)))
    _, s, vt = np.linalg.svd(K)
    free = vt[-1]  # Qdot of the free motion
    assert s[-1] > 1e-8 or K.shape[0] < K.shape[1]
    # Qdot_slot = omega x Q_slot for the three direction slots u, v = rp - rd, w: least squares for omega
    A, b = [], []
    for vec, dvec in ((Q0[0:3], free[0:3]), (Q0[3:6] - Q0[6:9], free[3:6] - free[6:9]), (Q0[9:12], free[9:12])):
        A.append(-np.array([[0, -vec[2], vec[1]], [vec[2], 0, -vec[0]], [-vec[1], vec[0], 0]]))  # omega x vec = -[vec]x omega
        b.append(dvec)
    omega = np.linalg.lstsq(np.vstack(A), np.concatenate(b), rcond=None)[0]
    axis = omega / np.linalg.norm(omega)
    rng = np.random.default_rng(seed)
    R = st.rotation_from_rotvec(rng.uniform(-np.pi, np.pi, (count, 1)) * axis)
    rate = rng.normal(0.0, 1.0, (count, 1)) * axis
    Q = np.concatenate([np.einsum("bij,j->bi", R, Q0[3 * k : 3 * k + 3]) for k in range(4)], axis=1)
    Qd = np.concatenate([np.cross(rate, Q[:, 3 * k : 3 * k + 3]) for k in range(4)], axis=1)
    return Q, Qd


def config_states(name, model, table, count, seed):
    if name == "pendulum":
        return hinge_swings(model, np.array([0, 0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0], float), count, seed)
    return rm.config_states(name, table, count, seed)


@pytest.fixture(scope="module", autouse=True)
def _error_table():
    yield
    out = os.path.join(REPO, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "live_reference_errors.json"), "w") as f:
        json.dump(ERRORS, f, indent=1, sort_keys=True)
    print("\nunfloored errors max|d|/max|ref| against the live reference:")
    for k, v in sorted(ERRORS.items()):
        print(f"  {k:36s} " + "  ".join(f"{kk} {vv:.2e}" for kk, vv in v.items()))


@pytest.mark.skipif(not rm.available(), reason="baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
@pytest.mark.parametrize("name", ["pendulum", "pendulum_force", "lower_limb", "knee", "full_body"])
def test_drop_in_matches_the_live_reference(name):
    warnings.filterwarnings("ignore")
    rm.import_reference()
    from bionc.bionc_numpy import NaturalCoordinates, NaturalVelocities

    from bionc_b200.bionc_cuda import BiomechanicalModel
    from bionc_b200.bionc_cuda.model_table import ModelTable

    ref_model, ref_fext = rm.CONFIGS[name]["build"]()  # .nmod / .nc loaded, or built with the reference's own API
    model = BiomechanicalModel.from_numpy(ref_model)  # reference object -> model table -> device
    assert model.nb_Q == ref_model.nb_Q and model.nb_holonomic_constraints == ref_model.nb_holonomic_constraints
    table = ModelTable.from_numpy(ref_model)
    Q, Qd = config_states(name, ref_model, table, N_EVAL, seed=11)

    # ---- one evaluation, 256 states; the force set is handed over as the reference's own object
    qdd, lam, status = model.forward_dynamics(Q, Qd, external_forces=ref_fext, return_status=True)
    assert not status.any()
    worst = dict(Qddot=0.0, lam=0.0, Qddot_unfloored=0.0, lam_unfloored=0.0)
    for s in range(N_EVAL):
        rq, rl = ref_model.forward_dynamics(NaturalCoordinates(Q[s]), NaturalVelocities(Qd[s]), external_forces=ref_fext)
        rq, rl = np.asarray(rq).reshape(-1), np.asarray(rl).reshape(-1)
        worst["Qddot"] = max(worst["Qddot"], rel_err(qdd[s], rq))
        worst["lam"] = max(worst["lam"], rel_err(lam[s], rl))
        worst["Qddot_unfloored"] = max(worst["Qddot_unfloored"], unfloored(qdd[s], rq))
        worst["lam_unfloored"] = max(worst["lam_unfloored"], unfloored(lam[s], rl))
    ERRORS[f"{name}/forward_dynamics"] = worst
    assert worst["Qddot"] < TOL_EVAL and worst["lam"] < TOL_EVAL, worst

    # ---- with Baumgarte stabilisation (biomechanical_model.py:416-421), 32 states
    stab = dict(alpha=0.7, beta=1.3)
    qdd, lam = model.forward_dynamics(Q[:32], Qd[:32], external_forces=ref_fext, stabilization=stab)
    ws = dict(Qddot=0.0, lam=0.0, Qddot_unfloored=0.0, lam_unfloored=0.0)
    for s in range(32):
        rq, rl = ref_model.forward_dynamics(NaturalCoordinates(Q[s]), NaturalVelocities(Qd[s]), external_forces=ref_fext, stabilization=stab)
        rq, rl = np.asarray(rq).reshape(-1), np.asarray(rl).reshape(-1)
        ws["Qddot"], ws["lam"] = max(ws["Qddot"], rel_err(qdd[s], rq)), max(ws["lam"], rel_err(lam[s], rl))
        ws["Qddot_unfloored"], ws["lam_unfloored"] = max(ws["Qddot_unfloored"], unfloored(qdd[s], rq)), max(ws["lam_unfloored"], unfloored(lam[s], rl))
    ERRORS[f"{name}/forward_dynamics_stabilized"] = ws
    assert ws["Qddot"] < TOL_EVAL and ws["lam"] < TOL_EVAL, ws

    # ---- 100 RK4 steps, 8 states, the reference's own integrator with normalize_idx = model.normalized_coordinates
    h = rm.CONFIGS[name]["h"]
    t = np.arange(TRAJ_STEPS + 1) * h
    Qf, Qdf, st_f = model.rk4(Q[:N_TRAJ], Qd[:N_TRAJ], t=t, normalize=True, external_forces=ref_fext, return_status=True)
    assert not st_f.any()
    wt = dict(state=0.0, state_unfloored=0.0)
    for s in range(N_TRAJ):
        rq, rqd = rm.reference_rk4(ref_model, Q[s], Qd[s], t, True, ref_fext)
        got, ref = np.concatenate([Qf[s], Qdf[s]]), np.concatenate([rq, rqd])
        wt["state"], wt["state_unfloored"] = max(wt["state"], rel_err(got, ref)), max(wt["state_unfloored"], unfloored(got, ref))
    ERRORS[f"{name}/rk4_{TRAJ_STEPS}_steps"] = wt
    assert wt["state"] < TOL_TRAJ, wt
