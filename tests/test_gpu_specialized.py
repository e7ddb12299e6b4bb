"""Model-specialised integrator kernels (bionc_cuda_model_specialize: the model tables as compile-time constants, NVRTC,
driver-API launch) against (a) the golden trajectories of the unmodified reference, at the tolerance of the generic
kernel's own test, and (b) the generic kernel on the same inputs -- the two run the same arithmetic in the same order,
so they must agree to rounding (1e-12 on the scale of the state; bitwise on every model measured so far)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import golden_case_names  # noqa: E402
from helpers import load_case, model_path, nvrtc_available, rel_err  # noqa: E402
from test_gpu_parity import TOL_TRAJ, _fext, _model, _states, _traj_tags  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not nvrtc_available(), reason="libnvrtc not installed: the generic kernels serve every call")]

TOL_SAME = 1e-12
TRAJ_CASES = [n for n in golden_case_names() if _traj_tags(load_case(n))]


def _pair(name):
    """(specialised-to-be, generic) handles of one model"""
    return _model(name), _model(name)


@pytest.mark.parametrize("name", TRAJ_CASES)
def test_specialized_rk4_matches_reference_and_generic_kernel(name):
    case = load_case(name)
    spec, gen = _pair(name)
    fext = _fext(spec, case)
    assert spec.specialized_kernels() == 0
    seconds = spec.specialize(external_forces=fext)
    assert seconds > 0.0 and spec.specialized_kernels() == 1
    assert spec.specialize(external_forces=fext) == 0.0  # that shape is compiled: nothing to do
    for tag in _traj_tags(case):
        t = case[f"traj_{tag}_t"]
        ref = case[f"traj_{tag}_final"]
        ns = ref.shape[0]
        kw = dict(t=t, normalize=bool(case[f"traj_{tag}_normalize"]), external_forces=fext, return_status=True)
        launches = spec.launch_count()
        Qf, Qdf, status = spec.rk4(case["Q"][:ns], case["Qdot"][:ns], **kw)
        assert spec.launch_count() == launches + 1
        Qg, Qdg, status_g = gen.rk4(case["Q"][:ns], case["Qdot"][:ns], **kw)
        assert not status.any() and not status_g.any()
        for s in range(ns):
            assert rel_err(np.concatenate([Qf[s], Qdf[s]]), ref[s]) < TOL_TRAJ, f"{name}/{tag} system {s} against the reference"
            assert rel_err(np.concatenate([Qf[s], Qdf[s]]), np.concatenate([Qg[s], Qdg[s]])) < TOL_SAME, f"{name}/{tag} system {s} against the generic kernel"


def test_specialized_shapes_and_fallback():
    """One kernel per call shape; a call of another shape runs the generic kernel of the same handle."""
    case = load_case("lower_limb")
    spec, gen = _pair("lower_limb")
    Q, Qd = _states("tree", spec, case, 96, seed=5)
    stab = dict(alpha=50.0, beta=20.0)
    assert spec.specialize(stabilization=stab) > 0.0  # lean + Baumgarte
    a = spec.rk4(Q, Qd, h=0.01, n_steps=50, stabilization=stab, return_status=True)
    b = gen.rk4(Q, Qd, h=0.01, n_steps=50, stabilization=stab, return_status=True)
    assert rel_err(np.concatenate([a[0], a[1]], axis=1), np.concatenate([b[0], b[1]], axis=1)) < TOL_SAME
    assert np.array_equal(a[2], b[2])
    # the unstabilised shape was not compiled: generic kernel, identical to the other handle bit for bit
    a = spec.rk4(Q, Qd, h=0.01, n_steps=20)
    b = gen.rk4(Q, Qd, h=0.01, n_steps=20)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert spec.specialize() > 0.0 and spec.specialized_kernels() == 2
    a = spec.rk4(Q, Qd, h=0.01, n_steps=20)
    assert rel_err(np.concatenate([a[0], a[1]], axis=1), np.concatenate([b[0], b[1]], axis=1)) < TOL_SAME


def test_specialized_per_system_inertia_and_force_values():
    # config 5: per-system inertial parameters (16 systems would do; the ragged last CTA is part of the test)
    case = load_case("full_body_perturbed")
    spec, gen = _pair("full_body")
    J = case["seg_J"]
    ip = np.concatenate([case["seg_mass"][..., None], case["seg_com"],
                         np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1)
    nt = case["traj_cfg_final"].shape[0]
    assert spec.specialize(per_system_inertia=True) > 0.0
    Qf, Qdf = spec.rk4(case["Q"][:nt], case["Qdot"][:nt], t=case["traj_cfg_t"], normalize=True, inertial_parameters=ip[:nt])
    Qg, Qdg = gen.rk4(case["Q"][:nt], case["Qdot"][:nt], t=case["traj_cfg_t"], normalize=True, inertial_parameters=ip[:nt])
    for s in range(nt):
        assert rel_err(np.concatenate([Qf[s], Qdf[s]]), case["traj_cfg_final"][s]) < TOL_TRAJ
        assert rel_err(np.concatenate([Qf[s], Qdf[s]]), np.concatenate([Qg[s], Qdg[s]])) < TOL_SAME
    # per-system, per-step force values on the pendulum
    case = load_case("pendulum_force_batch")
    spec, gen = _pair("pendulum_force_batch")
    fext = _fext(spec, case)
    B, steps = 70, 12
    Q, Qd = _states("hinge", spec, case, B, seed=9)
    F = int(case["fext_seg"].size)
    vals = np.random.default_rng(3).normal(size=(steps, B, F, 6))
    assert spec.specialize(external_forces=fext) > 0.0
    a = spec.rk4(Q, Qd, h=1.0 / 60.0, n_steps=steps, external_forces=fext, external_force_values=vals)
    b = gen.rk4(Q, Qd, h=1.0 / 60.0, n_steps=steps, external_forces=fext, external_force_values=vals)
    assert rel_err(np.concatenate([a[0], a[1]], axis=1), np.concatenate([b[0], b[1]], axis=1)) < TOL_SAME


def test_specialized_host_buffer_entry_point():
    case = load_case("knee_feikes")
    spec, gen = _pair("knee_feikes")
    Q, Qd = _states("knee", spec, case, 300, seed=2)
    assert spec.specialize() > 0.0
    Qa, Qda, Qb, Qdb = Q.copy(), Qd.copy(), Q.copy(), Qd.copy()
    spec.rk4_host(Qa, Qda, 5e-5, 40)
    gen.rk4_host(Qb, Qdb, 5e-5, 40)
    assert rel_err(np.concatenate([Qa, Qda], axis=1), np.concatenate([Qb, Qdb], axis=1)) < TOL_SAME


def test_rk4_specialises_large_general_path_calls_by_itself():
    case = load_case("pendulum_force_batch")
    spec, gen = _pair("pendulum_force_batch")
    fext = _fext(spec, case)
    Q, Qd = _states("hinge", spec, case, 20000, seed=4)
    small = spec.rk4(Q[:100], Qd[:100], h=1.0 / 60.0, n_steps=100, external_forces=fext)
    assert spec.specialized_kernels() == 0  # 1e4 system-steps: not worth a compilation
    a = spec.rk4(Q, Qd, h=1.0 / 60.0, n_steps=100, external_forces=fext)  # 2e6 system-steps of a general-path shape
    assert spec.specialized_kernels() == 1
    b = gen.rk4(Q, Qd, h=1.0 / 60.0, n_steps=100, external_forces=fext, specialize=False)
    assert gen.specialized_kernels() == 0
    assert rel_err(np.concatenate([a[0], a[1]], axis=1), np.concatenate([b[0], b[1]], axis=1)) < TOL_SAME
    assert np.array_equal(small[0], b[0][:100])
    # the lean path keeps the generic kernel
    lean = _model("lower_limb")
    Ql, Qdl = _states("tree", lean, load_case("lower_limb"), 20000, seed=4)
    lean.rk4(Ql, Qdl, h=0.01, n_steps=100)
    assert lean.specialized_kernels() == 0
