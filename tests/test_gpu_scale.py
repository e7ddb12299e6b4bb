"""Parity at BASELINE.json's full sizes (1e6 systems) through size-independent properties:
the KKT residual of every system, bitwise batch-position independence, equivariance under rotations
about the gravity axis, and invariants of the fused integrator."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import load_case, model_path  # noqa: E402

B_FULL = 1_000_000


def _model(name):
    from bionc_b200.bionc_cuda import BiomechanicalModel

    return BiomechanicalModel.load(model_path(name))


def _tiled_states(kind, model, case, B, seed):
    """B consistent states on the device: a pool of 2^15 seeded states tiled (generation is numpy)."""
    from bionc_b200.utils import states as st

    pool = min(B, 1 << 15)
    if kind == "tree":
        Q, Qd = st.tree_states(model.table, pool, seed=seed)
    elif kind == "hinge":
        Q, Qd = st.hinge_pendulum_states(pool, seed=seed)
    else:
        Q, Qd = st.rigidly_moved_states(case["Q"][0], pool, seed=seed, angle_max=0.3, omega_sigma=0.5)
    reps = (B + pool - 1) // pool
    Qt = torch.from_numpy(Q).cuda().repeat(reps, 1)[:B].contiguous()
    Qdt = torch.from_numpy(Qd).cuda().repeat(reps, 1)[:B].contiguous()
    return Qt, Qdt


def _fext(model, case):
    from bionc_b200.bionc_cuda import ExternalForceSet

    if case["fext_seg"].size == 0:
        return None
    fs = ExternalForceSet.empty_from_nb_segment(model.nb_segments)
    for s, k, p in zip(case["fext_seg"], case["fext_kind"], case["fext_param"]):
        fs.external_forces[int(s)].append((int(k), np.array(p)))
    return fs


@pytest.mark.parametrize("name,kind", [("lower_limb", "tree"), ("pendulum_force_batch", "hinge"), ("knee_feikes", "knee")])
def test_kkt_residual_of_one_million_systems(name, kind):
    """[G K^T; K 0][Qdd; lam] = [f; -bias] for EVERY system of a 1e6 batch (biomechanical_model.py:404-426).
    K, bias and f_ext come from the component kernels (themselves pinned against the reference goldens)."""
    case = load_case(name)
    model = _model(name)
    fext = _fext(model, case)
    Q, Qd = _tiled_states(kind, model, case, B_FULL, seed=11)
    qdd, lam, status = model.forward_dynamics(Q, Qd, external_forces=fext, return_status=True)
    assert int(status.count_nonzero().item()) == 0
    G = torch.from_numpy(model.mass_matrix).cuda()
    fg = torch.from_numpy(model.gravity_forces()[:, 0]).cuda()
    worst_dyn = worst_con = 0.0
    chunk = 100_000
    for lo in range(0, B_FULL, chunk):
        sl = slice(lo, lo + chunk)
        K = model.holonomic_constraints_jacobian(Q[sl])
        bias = model.holonomic_constraints_acceleration_bias(Qd[sl])
        f = fg.expand(chunk, -1)
        if fext is not None:
            f = f + model.natural_external_forces(Q[sl], fext)
        r_dyn = qdd[sl] @ G.T + torch.bmm(K.transpose(1, 2), lam[sl].unsqueeze(-1)).squeeze(-1) - f
        r_con = torch.bmm(K, qdd[sl].unsqueeze(-1)).squeeze(-1) + bias
        scale = torch.clamp(torch.maximum(lam[sl].abs().amax(dim=1), qdd[sl].abs().amax(dim=1)), min=1.0)
        worst_dyn = max(worst_dyn, float((r_dyn.abs().amax(dim=1) / scale).max().item()))
        worst_con = max(worst_con, float((r_con.abs().amax(dim=1) / scale).max().item()))
    # residuals are backward errors: the knee KKT has cond ~ 1e6, so 1e-10 on the residual is far
    # stricter than 1e-10 on the solution
    assert worst_dyn < 1e-10 and worst_con < 1e-10, (worst_dyn, worst_con)


def test_results_do_not_depend_on_batch_position_or_size():
    """A system's result is bit-identical wherever it sits in the batch (the kernel has no cross-system
    arithmetic and no atomics)."""
    case = load_case("lower_limb")
    model = _model("lower_limb")
    Q, Qd = _tiled_states("tree", model, case, 200_003, seed=12)
    qdd, lam = model.forward_dynamics(Q, Qd)
    idx = torch.randperm(Q.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))[:4099]
    qdd2, lam2 = model.forward_dynamics(Q[idx].contiguous(), Qd[idx].contiguous())
    assert torch.equal(qdd[idx], qdd2) and torch.equal(lam[idx], lam2)
    Qf, Qdf = model.rk4(Q, Qd, h=0.01, n_steps=7)
    Qf2, Qdf2 = model.rk4(Q[idx].contiguous(), Qd[idx].contiguous(), h=0.01, n_steps=7)
    assert torch.equal(Qf[idx], Qf2) and torch.equal(Qdf[idx], Qdf2)


@pytest.mark.parametrize("name,kind", [("lower_limb", "tree"), ("knee_feikes", "knee"), ("full_body", "tree"), ("n_link_4", "chain")])
def test_repeated_runs_are_bit_identical(name, kind):
    """The warps of a CTA hand data to each other through shared memory with few barriers; any missing
    ordering would show up as run-to-run differences.  Three runs of the same 20-step integration (and one
    with a different batch size, i.e. a different CTA/SM schedule) must agree bit for bit."""
    from bionc_b200.utils import states as st

    case = load_case(name)
    model = _model(name)
    B = 150_001
    if kind == "chain":
        Qn, Qdn = st.planar_chain_states(model.table, 1 << 14, seed=31)
        Q = torch.from_numpy(Qn).cuda().repeat(10, 1)[:B].contiguous()
        Qd = torch.from_numpy(Qdn).cuda().repeat(10, 1)[:B].contiguous()
    else:
        Q, Qd = _tiled_states(kind, model, case, B, seed=31)
    h = 5e-5 if kind == "knee" else 1e-3
    ref = model.rk4(Q, Qd, h=h, n_steps=20, return_lambdas=True)
    for _ in range(2):
        out = model.rk4(Q, Qd, h=h, n_steps=20, return_lambdas=True)
        assert all(torch.equal(a, b) for a, b in zip(ref, out))
    sub = model.rk4(Q[:40_003].contiguous(), Qd[:40_003].contiguous(), h=h, n_steps=20, return_lambdas=True)
    assert all(torch.equal(a[:40_003], b) for a, b in zip(ref, sub))


def _rotate_about_z(X, ang):
    """Rotate every 3-vector of (B, 12N) states by `ang` (B,) about the global z axis (the gravity axis)."""
    B, n = X.shape
    v = X.reshape(B, n // 3, 3)
    c, s = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    out = torch.stack([c * v[..., 0] - s * v[..., 1], s * v[..., 0] + c * v[..., 1], v[..., 2]], dim=-1)
    return out.reshape(B, n)


def test_equivariance_under_rotation_about_the_gravity_axis():
    """The free-floating lower limb has no preferred horizontal direction: rotating the state about z rotates
    Qddot and leaves the rigid-body multipliers unchanged (joint multipliers rotate as vectors)."""
    case = load_case("lower_limb")
    model = _model("lower_limb")
    B = 300_000
    Q, Qd = _tiled_states("tree", model, case, B, seed=13)
    ang = torch.rand(B, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 6.0
    qdd, lam = model.forward_dynamics(Q, Qd)
    qdd_r, lam_r = model.forward_dynamics(_rotate_about_z(Q, ang), _rotate_about_z(Qd, ang))
    scale = torch.clamp(qdd.abs().amax(dim=1, keepdim=True), min=1.0)
    assert float(((_rotate_about_z(qdd, ang) - qdd_r).abs() / scale).max().item()) < 1e-10
    rigid_rows = torch.tensor([r + k for r in model.table.seg_rigid_row for k in range(6)], device="cuda")
    lscale = torch.clamp(lam.abs().amax(dim=1, keepdim=True), min=1.0)
    assert float(((lam[:, rigid_rows] - lam_r[:, rigid_rows]).abs() / lscale).max().item()) < 1e-10


def test_fused_rk4_invariants_at_full_batch():
    """1e6 systems x 100 steps: u and w stay unit vectors (post-step renormalisation), nothing is flagged,
    the mechanical energy of the unforced free-floating chain drifts only at the integrator's order, and
    10 launches of 10 steps equal one launch of 100 steps bit for bit."""
    case = load_case("lower_limb")
    model = _model("lower_limb")
    Q, Qd = _tiled_states("tree", model, case, B_FULL, seed=14)
    Qa, Qda, status = model.rk4(Q, Qd, h=0.001, n_steps=100, normalize=True, return_status=True)
    assert int(status.count_nonzero().item()) == 0
    v = Qa.reshape(B_FULL, -1, 4, 3)
    assert float((v[:, :, 0].norm(dim=-1) - 1).abs().max().item()) < 1e-14
    assert float((v[:, :, 3].norm(dim=-1) - 1).abs().max().item()) < 1e-14
    Qb, Qdb = Q.clone(), Qd.clone()
    for _ in range(10):
        Qb, Qdb = model.rk4(Qb, Qdb, h=0.001, n_steps=10, normalize=True)
    assert torch.equal(Qa, Qb) and torch.equal(Qda, Qdb)
    # energy: E = 1/2 Qd^T G Qd - f_g^T Q (gravity is a constant natural force)
    G = torch.from_numpy(model.mass_matrix).cuda()
    fg = torch.from_numpy(model.gravity_forces()[:, 0]).cuda()
    E0 = 0.5 * ((Qd @ G) * Qd).sum(dim=1) - Q @ fg
    E1 = 0.5 * ((Qda @ G) * Qda).sum(dim=1) - Qa @ fg
    rel = ((E1 - E0).abs() / torch.clamp(E0.abs(), min=1.0)).max().item()
    assert rel < 1e-6, rel


def test_knee_parallel_mechanism_1000_steps_full_batch():
    """Config 4: closed-loop knee, 1e6 systems at rest in rigidly rotated poses, 1000 steps at
    h = 5e-5 (the reference itself diverges beyond t ~ 0.1 s, SURVEY finding 3); nothing may be flagged and
    the loop-closure constraints must stay satisfied at the drift level."""
    from bionc_b200.utils import states as st

    case = load_case("knee_feikes")
    model = _model("knee_feikes")
    pool = 1 << 15
    Qp, Qdp = st.rigidly_moved_states(case["Q"][0], pool, seed=15, angle_max=0.3, omega_sigma=0.0)
    reps = (B_FULL + pool - 1) // pool
    Q = torch.from_numpy(Qp).cuda().repeat(reps, 1)[:B_FULL].contiguous()
    Qd = torch.from_numpy(Qdp).cuda().repeat(reps, 1)[:B_FULL].contiguous()
    Qf, Qdf, status = model.rk4(Q, Qd, h=5e-5, n_steps=1000, normalize=True, return_status=True)
    assert int(status.count_nonzero().item()) == 0
    phi = model.holonomic_constraints(Qf[:200_000])
    # unstabilised drift: the reference's own end states in the golden file sit at |Phi| = 1.5e-4 .. 7.6e-4
    assert float(phi.abs().max().item()) < 5e-3
    # three systems against the reference's own 1000-step end states (golden file)
    ref = case["traj_cfg_final"]
    ns = ref.shape[0]
    Qg, Qdg = model.rk4(case["Q"][:ns], case["Qdot"][:ns], t=case["traj_cfg_t"], normalize=True)
    err = np.abs(np.concatenate([Qg, Qdg], axis=1) - ref).max() / max(1.0, np.abs(ref).max())
    assert err < 1e-8


def test_status_flags_mirror_the_reference_linalg_error():
    """With initial angular velocities some knee trajectories hit the reference's instability inside the 1000
    steps: there the reference algorithm (oracle) raises LinAlgError / goes non-finite, and the kernel flags
    exactly such systems; unflagged ones still match the oracle to 1e-8."""
    from bionc_oracle import OracleModel
    from bionc_b200.utils import states as st

    case = load_case("knee_feikes")
    model = _model("knee_feikes")
    om = OracleModel(model_path("knee_feikes"))
    Q, Qd = st.rigidly_moved_states(case["Q"][0], 4096, seed=15, angle_max=0.3, omega_sigma=0.5)
    t = np.arange(1001) * 5e-5
    Qf, Qdf, status = model.rk4(Q, Qd, t=t, normalize=True, return_status=True)
    bad, good = np.nonzero(status)[0], np.nonzero(status == 0)[0]
    assert good.size > 0
    # a spread of the flagged and of the unflagged systems (an oracle trajectory is a second of numpy each)
    for s in bad[:: max(1, bad.size // 5)][:5]:
        with np.errstate(all="ignore"):
            try:
                rq, rqd = om.rk4(Q[s], Qd[s], t, normalize=True)
                failed = not (np.isfinite(rq).all() and np.isfinite(rqd).all()) or np.abs(rq).max() > 1e6
            except np.linalg.LinAlgError:
                failed = True
        assert failed, f"system {s} flagged by the kernel but fine in the oracle"
    for s in good[:: max(1, good.size // 5)][:5]:
        rq, rqd = om.rk4(Q[s], Qd[s], t, normalize=True)
        assert np.abs(np.concatenate([Qf[s] - rq, Qdf[s] - rqd])).max() < 1e-8


def test_pendulum_with_force_1000_steps_full_batch():
    """Config 2: pendulum_with_force.nmod with the `no_equilibrium` external force, 1e6 hinge states,
    h = 1/60, 1000 steps, normalize=True.  Every trajectory stays on the hinge (rp pinned, |Phi| at drift level,
    nothing flagged); a seeded subsample follows the oracle's (reference algorithm) 1000-step trajectory to 1e-8."""
    from bionc_oracle import OracleModel

    from helpers import fext_list

    case = load_case("pendulum_force_batch")
    model = _model("pendulum_force_batch")
    fext = _fext(model, case)
    Q, Qd = _tiled_states("hinge", model, case, B_FULL, seed=21)
    h, steps = 1.0 / 60.0, 1000
    Qf, Qdf, status = model.rk4(Q, Qd, h=h, n_steps=steps, normalize=True, external_forces=fext, return_status=True)
    assert int(status.count_nonzero().item()) == 0
    assert bool(torch.isfinite(Qf).all()) and bool(torch.isfinite(Qdf).all())
    norms = model.constraint_residual_norms(Qf)
    # unstabilised drift over 16.7 s of simulated time (worst system 3.9e-3; the oracle comparison below shows the
    # reference algorithm drifts identically): a bound on blow-ups, not an accuracy claim
    assert float(norms.max().item()) < 2e-2
    v = Qf.reshape(B_FULL, 1, 4, 3)
    assert float((v[:, :, 0].norm(dim=-1) - 1).abs().max().item()) < 1e-14
    om = OracleModel(model_path("pendulum_force_batch"))
    t = np.arange(steps + 1) * h
    for s in (0, 12345, 999_999):
        rq, rqd = om.rk4(Q[s].cpu().numpy(), Qd[s].cpu().numpy(), t, normalize=True, fext=fext_list(case))
        got = np.concatenate([Qf[s].cpu().numpy(), Qdf[s].cpu().numpy()])
        ref = np.concatenate([rq, rqd])
        # t is an arange here and a constant h on the device: the step sizes differ in the last bit
        assert np.abs(got - ref).max() / max(1.0, np.abs(ref).max()) < 1e-8, s


def test_full_body_perturbed_inertia_full_batch():
    """Config 5: the 10-segment full body with per-system natural inertial parameters (masses x U(0.9,
    1.1), COM x U(0.95, 1.05), inertias x U(0.9, 1.1)), 1.25e6 systems (the per-GPU share of 1e7 on 8 GPUs), 100
    steps.  Nothing flagged, unit vectors kept, and a subsample against the oracle rebuilt with that system's
    parameters (forward dynamics of every sampled system, one full 100-step trajectory)."""
    from bionc_oracle import OracleModel

    from bionc_b200.utils import states as st

    B = 1_250_000
    case = load_case("full_body")
    model = _model("full_body")
    Q, Qd = _tiled_states("tree", model, case, B, seed=22)
    pool = 1 << 12
    mass, com, J = st.perturbed_inertial_parameters(model.table, pool, seed=23)
    ip_pool = np.concatenate(
        [mass[..., None], com, np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1)
    reps = (B + pool - 1) // pool
    ip = torch.from_numpy(ip_pool).cuda().repeat(reps, 1, 1)[:B].contiguous()
    h, steps = 0.001, 100
    Qf, Qdf, status = model.rk4(Q, Qd, h=h, n_steps=steps, normalize=True, inertial_parameters=ip, return_status=True)
    assert int(status.count_nonzero().item()) == 0
    v = Qf.reshape(B, model.nb_segments, 4, 3)
    assert float((v[:, :, 0].norm(dim=-1) - 1).abs().max().item()) < 1e-14
    assert float(model.constraint_residual_norms(Qf).max().item()) < 1e-6
    qdd, lam = model.forward_dynamics(Q[:pool], Qd[:pool], inertial_parameters=ip[:pool])
    for s in (0, 7, 4095):
        om = OracleModel(model_path("full_body"), seg_mass=mass[s], seg_com=com[s], seg_J=J[s])
        rq, rl = om.forward_dynamics(Q[s].cpu().numpy(), Qd[s].cpu().numpy())
        assert np.abs(qdd[s].cpu().numpy() - rq).max() / max(1.0, np.abs(rq).max()) < 1e-10
        assert np.abs(lam[s].cpu().numpy() - rl).max() / max(1.0, np.abs(rl).max()) < 1e-10
    s = 4095 + pool  # a tiled copy: same state and parameters as system 4095
    om = OracleModel(model_path("full_body"), seg_mass=mass[4095], seg_com=com[4095], seg_J=J[4095])
    rq, rqd = om.rk4(Q[s].cpu().numpy(), Qd[s].cpu().numpy(), np.arange(steps + 1) * h, normalize=True)
    got = np.concatenate([Qf[s].cpu().numpy(), Qdf[s].cpu().numpy()])
    ref = np.concatenate([rq, rqd])
    assert np.abs(got - ref).max() / max(1.0, np.abs(ref).max()) < 1e-8


@pytest.mark.parametrize("name", ["lower_limb", "knee_feikes", "n_link_4", "full_body", "pendulum_force_batch", "tj_universal"])
def test_no_shared_memory_word_is_read_before_it_is_written(name):
    """bionc_cuda_debug_poison fills regions of a CTA's working set with NaN before the first evaluation.  A fresh CTA
    inherits whatever the previous one left in shared memory, so an evaluation must write every word before reading it:
    no output may change, for any region.  (Round 1 zeroed the right-hand side of the padding rows at the END of the
    joint-level system although the fill-reducing ordering moves the padded node: 0 x garbage, visible only as sporadic
    NaNs at large batches.)"""
    from bionc_b200.bionc_cuda import _lib

    lib = _lib.load()
    case, model = load_case(name), _model(name)
    Q, Qd = np.tile(case["Q"], (40, 1))[:67], np.tile(case["Qdot"], (40, 1))[:67]

    def run():
        return (model.forward_dynamics(Q, Qd, pivoting="never") + model.rk4(Q, Qd, h=1e-4, n_steps=3)
                + model.forward_dynamics(Q, Qd, pivoting="never", stabilization=dict(alpha=1.0, beta=1.0)))

    try:
        lib.bionc_cuda_debug_poison(0)
        ref = run()
        for bit in range(7):
            lib.bionc_cuda_debug_poison(1 << bit)
            out = run()
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, out)), f"region {bit}"
    finally:
        lib.bionc_cuda_debug_poison(0)


def test_forward_dynamics_repeats_bit_for_bit_at_scale():
    """The single-evaluation kernel on 150 001 hinge-chain systems (16 systems per CTA, level-parallel schedule, a padded
    node in the middle of the ordering), three times."""
    from bionc_b200.utils import states as st

    model = _model("n_link_4")
    Qn, Qdn = st.planar_chain_states(model.table, 1 << 14, seed=31)
    Q = torch.from_numpy(Qn).cuda().repeat(10, 1)[:150_001].contiguous()
    Qd = torch.from_numpy(Qdn).cuda().repeat(10, 1)[:150_001].contiguous()
    ref = model.forward_dynamics(Q, Qd, pivoting="never", return_status=True)
    assert not bool(torch.isnan(ref[0]).any()) and not bool(ref[2].any())
    for _ in range(3):
        out = model.forward_dynamics(Q, Qd, pivoting="never", return_status=True)
        assert all(torch.equal(a, b) for a, b in zip(ref, out))


@pytest.mark.parametrize("name,kind", [("lower_limb", "tree"), ("full_body", "tree")])
@pytest.mark.parametrize("B", [1, 31, 32, 33, 1000, 150_001])
def test_bulk_copy_forward_dynamics_matches_the_per_lane_kernel(name, kind, B):
    """forward_dynamics on the lean path moves whole system rows with the bulk-copy engine and evaluates from them
    (forward_dynamics_rows_kernel); bit 8 of bionc_cuda_debug_poison selects the per-lane load/store kernel instead.  Same
    arithmetic, so accelerations, multipliers and flags must agree bit for bit -- full tiles, partial tiles, one system."""
    from bionc_b200.bionc_cuda import _lib

    lib = _lib.load()
    case, model = load_case(name), _model(name)
    Q, Qd = _tiled_states(kind, model, case, B, seed=5)
    try:
        lib.bionc_cuda_debug_poison(0)
        rows = model.forward_dynamics(Q, Qd, pivoting="never", return_status=True)
        lib.bionc_cuda_debug_poison(0x100)
        lanes = model.forward_dynamics(Q, Qd, pivoting="never", return_status=True)
    finally:
        lib.bionc_cuda_debug_poison(0)
    assert not bool(torch.isnan(rows[0]).any()) and not bool(torch.isnan(rows[1]).any())
    assert all(torch.equal(a, b) for a, b in zip(rows, lanes))


def test_misaligned_device_arrays_are_refused_and_the_mirror_realigns():
    """The kernels use 16-byte accesses on the per-system arrays.  The C ABI refuses a pointer at an odd double (no launch,
    no CUDA fault that would poison the context); the Python mirror copies such a view to an aligned tensor."""
    import ctypes as C

    from bionc_b200.bionc_cuda import _lib

    lib = _lib.load()
    case, model = load_case("lower_limb"), _model("lower_limb")
    B, n = 64, model.nb_Q
    Q, Qd = _tiled_states("tree", model, case, B, seed=9)
    ref = model.forward_dynamics(Q, Qd, pivoting="never")
    buf = torch.zeros(B * n + 1, dtype=torch.float64, device="cuda")
    odd = buf[1:].view(B, n)
    odd.copy_(Q)
    assert odd.data_ptr() % 16 == 8 and odd.is_contiguous()
    out = model.forward_dynamics(odd, Qd, pivoting="never")
    assert all(torch.equal(a, b) for a, b in zip(ref, out))
    qdd = torch.empty_like(Q)
    rc = lib.bionc_cuda_forward_dynamics(model._h, C.c_void_p(odd.data_ptr()), C.c_void_p(Qd.data_ptr()), None,
                                         C.c_void_p(qdd.data_ptr()), None, None, B, None)
    assert rc != 0 and b"aligned" in lib.bionc_cuda_last_error()
    rc = lib.bionc_cuda_rk4(model._h, C.c_void_p(odd.data_ptr()), C.c_void_p(Qd.data_ptr()), 1e-3, None, 1, 1, None, None, 1, None,
                            None, B, None)
    assert rc != 0 and b"aligned" in lib.bionc_cuda_last_error()
    torch.cuda.synchronize()  # the context is intact
    assert all(torch.equal(a, b) for a, b in zip(ref, model.forward_dynamics(Q, Qd, pivoting="never")))
