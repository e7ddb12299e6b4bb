"""External forces as the reference has them: evaluated for every call of the dynamics
(ExternalForceSet.to_natural_external_forces(Q), external_force.py:143-180), so every trajectory may carry its own loads
and a dynamics closure may change them from step to step (utils/ode_solver.py:107-116).  Here: per-system values
(B, F, 6), per-step values (steps, B, F, 6), the host-buffer entry points, and the re-entrancy of the model handle
(the force table is a per-call object: two host threads drive ONE model with different force sets on their own streams)."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import fext_list, load_case, model_path, rel_err  # noqa: E402
from test_gpu_parity import _fext, _model  # noqa: E402

TOL_EVAL, TOL_TRAJ = 1e-10, 1e-8


def _per_system_lists(case, values):
    """Oracle force lists, one per system: the golden case's forces with system s' own (torque, force)."""
    out = []
    for s in range(values.shape[0]):
        lst = []
        for k, (seg, kind, par) in enumerate(fext_list(case)):
            p = np.array(par)
            p[0:6] = values[s, k]
            lst.append((seg, kind, p))
        out.append(lst)
    return out


@pytest.mark.parametrize("name", ["pendulum_force_allkinds", "double_pendulum_force_no_equilibrium"])
def test_per_system_force_values_match_the_oracle(name):
    from bionc_oracle import OracleModel

    case, model, om = load_case(name), _model(name), OracleModel(model_path(name))
    fext = _fext(model, case)
    F = case["fext_seg"].size
    rng = np.random.default_rng(3)
    B = 37
    Q = case["Q"][rng.integers(0, case["Q"].shape[0], B)]
    Qd = case["Qdot"][rng.integers(0, case["Q"].shape[0], B)] * rng.uniform(0.5, 1.5, (B, 1))
    vals = rng.normal(0.0, 2.0, (B, F, 6))
    lists = _per_system_lists(case, vals)
    fn = model.natural_external_forces(Q, fext, external_force_values=vals)
    qdd, lam, status = model.forward_dynamics(Q, Qd, external_forces=fext, external_force_values=vals, return_status=True)
    qdd_s, lam_s = model.forward_dynamics(Q, Qd, external_forces=fext, external_force_values=vals, stabilization=dict(alpha=0.7, beta=1.3))
    qdd_d, lam_d = model.forward_dynamics(Q, Qd, external_forces=fext, external_force_values=vals, pivoting="always")
    assert not status.any()
    for s in range(B):
        assert rel_err(fn[s], om.natural_external_forces(Q[s], lists[s])) < TOL_EVAL
        rq, rl = om.forward_dynamics(Q[s], Qd[s], lists[s])
        assert rel_err(qdd[s], rq) < TOL_EVAL and rel_err(lam[s], rl) < TOL_EVAL, s
        assert rel_err(qdd_d[s], rq) < TOL_EVAL and rel_err(lam_d[s], rl) < TOL_EVAL, s
        rq, rl = om.forward_dynamics(Q[s], Qd[s], lists[s], dict(alpha=0.7, beta=1.3))
        assert rel_err(qdd_s[s], rq) < TOL_EVAL and rel_err(lam_s[s], rl) < TOL_EVAL, s
    # shared values are the special case of equal per-system values
    same = np.broadcast_to(case["fext_param"][None, :, 0:6], (B, F, 6)).copy()
    a = model.forward_dynamics(Q, Qd, external_forces=fext, external_force_values=same)
    b = model.forward_dynamics(Q, Qd, external_forces=fext)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # host-buffer entry point with host values
    qh, lh, sh = model.forward_dynamics_host(Q, Qd, external_forces=fext, external_force_values=vals)
    assert np.array_equal(qh, qdd) and np.array_equal(lh, lam) and not sh.any()


def test_per_step_force_values_in_the_fused_integrator():
    """(steps, B, F, 6): step k of the launch uses slice k, held over the four stages -- the oracle integrates step by step
    with that step's force list.  Fewer slices than steps: the last one is held."""
    from bionc_oracle import OracleModel

    name = "pendulum_force_allkinds"
    case, model, om = load_case(name), _model(name), OracleModel(model_path(name))
    fext = _fext(model, case)
    F = case["fext_seg"].size
    rng = np.random.default_rng(5)
    B, steps, given = 6, 12, 9
    Q, Qd = case["Q"][:1].repeat(B, 0), case["Qdot"][:1].repeat(B, 0) * rng.uniform(0.5, 1.5, (B, 1))
    vals = rng.normal(0.0, 1.0, (given, B, F, 6))
    t = np.arange(steps + 1) * 0.01
    Qf, Qdf, traj, lam, status = model.rk4(Q, Qd, t=t, normalize=True, external_forces=fext, external_force_values=vals,
                                           save_every=4, return_lambdas=True, return_status=True)
    assert not status.any() and traj.shape == (3, B, 24)
    for s in range(B):
        q, qd = Q[s], Qd[s]
        for k in range(steps):
            lst = _per_system_lists(case, vals[min(k, given - 1)])[s]
            if k == steps - 1:
                assert rel_err(lam[s], om.forward_dynamics(q, qd, lst)[1]) < 1e-9  # multipliers of the first stage of the last step
            q, qd = om.rk4(q, qd, t[k : k + 2], normalize=True, fext=lst)
            if (k + 1) % 4 == 0:
                assert rel_err(traj[(k + 1) // 4 - 1, s], np.concatenate([q, qd])) < TOL_TRAJ
        assert rel_err(np.concatenate([Qf[s], Qdf[s]]), np.concatenate([q, qd])) < TOL_TRAJ
    # host-buffer entry point, values on the host
    Qh, Qdh = np.ascontiguousarray(Q.copy()), np.ascontiguousarray(Qd.copy())
    for k in range(steps):  # constant h: one call per step with that step's slice == one call with all slices
        pass
    model.rk4_host(Qh, Qdh, 0.01, steps, normalize=True, external_forces=fext, external_force_values=vals)
    Qc, Qdc = model.rk4(Q, Qd, h=0.01, n_steps=steps, normalize=True, external_forces=fext, external_force_values=vals)
    assert np.array_equal(Qh, Qc) and np.array_equal(Qdh, Qdc)


def test_one_model_two_threads_different_force_sets():
    """The model handle is immutable after creation and the force table is a per-call object uploaded on the caller's
    stream: two host threads with different force sets and their own streams get, bit for bit, what they get alone."""
    from bionc_b200.bionc_cuda import ExternalForceSet

    name = "pendulum_force_batch"
    case, model = load_case(name), _model(name)
    B = 20_000
    rng = np.random.default_rng(0)
    idx = rng.integers(0, case["Q"].shape[0], B)
    Q = torch.from_numpy(case["Q"][idx]).cuda()
    Qd = torch.from_numpy(case["Qdot"][idx] * rng.uniform(0.5, 1.5, (B, 1))).cuda()
    sets = []
    for k in range(2):
        fs = ExternalForceSet.empty_from_nb_segment(1)
        fs.add_in_global_local_point(0, np.array([0.1 * k, 0, 9.81 * (k + 1), 0, 1.0 - k, 0]), point_in_local=np.array([0, 0.25 + 0.1 * k, 0]))
        if k == 1:
            fs.add_in_global(0, np.array([0.3, 0, 0, 0, 0, 2.0]), point_in_global=np.array([0.1, 0.2, 0.3]))
        sets.append(fs)
    alone = [model.rk4(Q, Qd, h=1e-3, n_steps=50, external_forces=fs) for fs in sets]
    alone_fd = [model.forward_dynamics(Q, Qd, external_forces=fs, pivoting="never") for fs in sets]
    torch.cuda.synchronize()
    results, errors = [None, None], []

    def worker(k):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                out = []
                for _ in range(8):  # several calls in flight per thread
                    out = [model.rk4(Q, Qd, h=1e-3, n_steps=50, external_forces=sets[k]),
                           model.forward_dynamics(Q, Qd, external_forces=sets[k], pivoting="never")]
            st.synchronize()
            results[k] = out
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for k in range(2):
        assert torch.equal(results[k][0][0], alone[k][0]) and torch.equal(results[k][0][1], alone[k][1])
        assert torch.equal(results[k][1][0], alone_fd[k][0]) and torch.equal(results[k][1][1], alone_fd[k][1])
    assert not torch.equal(alone[0][0], alone[1][0])  # the two force sets do give different trajectories


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Every entry point switches to the model's device and back (the caller's current device is left alone), and the
    shared-memory opt-in of the kernels is taken per device."""
    from bionc_b200.bionc_cuda import BiomechanicalModel
    from bionc_b200.utils import states as st

    m0 = BiomechanicalModel.load(model_path("full_body"), device="cuda:0")
    m1 = BiomechanicalModel.load(model_path("full_body"), device="cuda:1")  # > 48 KB of dynamic shared memory on both
    Q, Qd = st.tree_states(m0.table, 64, seed=0)
    torch.cuda.set_device(0)
    a = m1.forward_dynamics(torch.from_numpy(Q).to("cuda:1"), torch.from_numpy(Qd).to("cuda:1"))
    assert torch.cuda.current_device() == 0
    b = m0.forward_dynamics(torch.from_numpy(Q).to("cuda:0"), torch.from_numpy(Qd).to("cuda:0"))
    assert torch.equal(a[0].cpu(), b[0].cpu()) and torch.equal(a[1].cpu(), b[1].cpu())
    ra = m1.rk4(Q, Qd, h=1e-3, n_steps=5)
    rb = m0.rk4(Q, Qd, h=1e-3, n_steps=5)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])


def test_joint_generalized_forces_behind_the_flag():
    """SURVEY 8(f) rank 3.  Default: accepted and ignored exactly like the reference (call site commented out,
    biomechanical_model.py:391-413).  With apply_joint_generalized_forces=True: the documented semantics of
    generalized_force.py:98-129, against vectors composed from the reference's own existing methods
    (tests/golden/generate_joint_forces.py -- the reference's to_natural_joint_forces itself raises AttributeError)."""
    import os

    from helpers import GOLDEN
    from bionc_b200.bionc_cuda import BiomechanicalModel, ExternalForceSet
    from bionc_b200.bionc_cuda.generalized_force import JointGeneralizedForces, JointGeneralizedForcesList
    from bionc_b200.bionc_cuda.model_table import ModelTable

    with np.load(os.path.join(GOLDEN, "cases_extra", "joint_forces_lower_limb.npz")) as z:
        g = {k: np.array(z[k]) for k in z.files}
    table = ModelTable.load(model_path("lower_limb"))
    table.all_joint_names = [f"j{k}" for k in range(g["all_joint_parent"].size)]  # the committed table predates format 3
    table.all_joint_parent, table.all_joint_child = g["all_joint_parent"], g["all_joint_child"]
    model = BiomechanicalModel.from_table(table)
    jl = JointGeneralizedForcesList.empty_from_nb_joint(len(table.all_joint_names))
    for j, v in zip(g["joint_index"], g["joint_force"]):
        jl.add_generalized_force(int(j), JointGeneralizedForces(v, np.zeros(3)))
    Q, Qd = g["Q"], g["Qdot"]
    plain = model.forward_dynamics(Q, Qd)
    ignored = model.forward_dynamics(Q, Qd, joint_generalized_forces=jl)
    assert np.array_equal(plain[0], ignored[0]) and np.array_equal(plain[1], ignored[1])  # the reference's behaviour
    qdd, lam, status = model.forward_dynamics(Q, Qd, joint_generalized_forces=jl, apply_joint_generalized_forces=True, return_status=True)
    assert not status.any()
    for s in range(Q.shape[0]):
        assert rel_err(qdd[s], g["ref_Qddot"][s]) < TOL_EVAL and rel_err(lam[s], g["ref_lam"][s]) < TOL_EVAL, s
    assert rel_err(qdd, plain[0]) > 1e-3  # and they do act
    # the natural joint forces themselves, through the force evaluator
    seg, kind, par = jl.to_arrays(table.all_joint_parent, table.all_joint_child)
    fs = ExternalForceSet.empty_from_nb_segment(model.nb_segments)
    for s_, k_, p_ in zip(seg, kind, par):
        fs.external_forces[int(s_)].append((int(k_), p_))
    fn = model.natural_external_forces(Q, fs)
    assert rel_err(fn, g["ref_natural_joint_forces"]) < 1e-12
    # fused integrator and dense kernel take the same table
    qd2, lam2 = model.forward_dynamics(Q, Qd, joint_generalized_forces=jl, apply_joint_generalized_forces=True, pivoting="always")
    assert rel_err(qd2, qdd) < 1e-9 and rel_err(lam2, lam) < 1e-9
    from bionc_oracle import OracleModel

    om = OracleModel(model_path("lower_limb"))
    fext = [(int(a), int(b), c) for a, b, c in zip(seg, kind, par)]
    t = np.linspace(0.0, 0.05, 26)
    Qf, Qdf = model.rk4(Q[:2], Qd[:2], t=t, joint_generalized_forces=jl, apply_joint_generalized_forces=True)
    for s in range(2):
        rq, rqd = om.rk4(Q[s], Qd[s], t, normalize=True, fext=fext)
        assert rel_err(np.concatenate([Qf[s], Qdf[s]]), np.concatenate([rq, rqd])) < TOL_TRAJ


def test_force_table_does_not_reference_host_memory_after_the_call():
    """The per-call force table is written to the device by a parameter-carrying kernel.  A call queued behind a long
    kernel on a busy stream (where an asynchronous copy from pageable memory would read its source late) gives the
    same bits as on an idle stream, and the force description may be dropped as soon as the call returns."""
    name = "pendulum_force_batch"
    case, model = load_case(name), _model(name)
    B = 300_000
    rng = np.random.default_rng(2)
    idx = rng.integers(0, case["Q"].shape[0], B)
    Q, Qd = torch.from_numpy(case["Q"][idx]).cuda(), torch.from_numpy(case["Qdot"][idx]).cuda()
    fext = _fext(model, case)
    ref = model.forward_dynamics(Q[:4096], Qd[:4096], external_forces=fext, pivoting="never")
    torch.cuda.synchronize()
    outs = []
    for _ in range(3):
        model.rk4(Q, Qd, h=1e-3, n_steps=300, external_forces=fext)  # tens of milliseconds of queued work
        outs.append(model.forward_dynamics(Q[:4096], Qd[:4096], external_forces=_fext(model, case), pivoting="never"))
    torch.cuda.synchronize()
    for out in outs:
        assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
