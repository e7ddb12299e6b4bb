"""
Golden-vector generator.  Runs ONLY in the build container, where the unmodified reference lives
at /root/reference; it imports the real ``bionc`` package through the import shims in
``oracle/stubs`` (casadi / ezc3d / pyomeca / trc / biorbd are absent here and only needed for
type hints), evaluates the reference on seeded inputs and writes

* ``tests/golden/models/<name>.npz`` : the compiled model table (``ModelTable.from_numpy``)
* ``tests/golden/cases/<name>.npz``  : inputs and the reference's outputs
  (``Phi_r, Phi_j, K_j, Phi_h, K_h, bias, f_ext, Qddot, lambda``, Baumgarte variants, RK4 end states)

Usage (from a scratch directory -- the reference examples write ``.nmod`` files into the CWD):

    cd $(mktemp -d) && python /root/repo/tests/golden/generate_golden.py

Nothing here runs on the GPU box; the committed ``.npz`` files are the fixtures.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import time
import warnings

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = os.environ.get("BIONC_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REPO, "oracle", "stubs"), REF, REPO]
warnings.filterwarnings("ignore", category=SyntaxWarning)

import bionc  # noqa: E402  (the unmodified reference)
from bionc import CartesianAxis, JointType, NaturalAxis, TransformationMatrixType  # noqa: E402
from bionc.bionc_numpy import (  # noqa: E402
    BiomechanicalModel,
    ExternalForceSet,
    NaturalCoordinates,
    NaturalSegment,
    NaturalVelocities,
    SegmentNaturalCoordinates,
)
from bionc.utils.ode_solver import RK4  # noqa: E402

from bionc_b200.bionc_cuda.model_table import (  # noqa: E402
    FEXT_GLOBAL,
    FEXT_GLOBAL_LOCAL_POINT,
    FEXT_GLOBAL_ON_PROXIMAL,
    FEXT_LOCAL,
    FEXT_NPARAM,
    ModelTable,
)
from bionc_b200.utils import states as st  # noqa: E402

OUT_MODELS = os.path.join(REPO, "tests", "golden", "models")
OUT_CASES = os.path.join(REPO, "tests", "golden", "cases")


def load_example(relpath: str):
    path = os.path.join(REF, relpath)
    d = os.path.dirname(path)
    if d not in sys.path:
        sys.path.insert(0, d)
    spec = importlib.util.spec_from_file_location(os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------- external forces
def fext_to_arrays(fext: ExternalForceSet | None):
    """Flatten a reference ExternalForceSet to (segment, kind, param24) arrays."""
    seg, kind, par = [], [], []
    if fext is not None:
        for s, forces in enumerate(fext.external_forces):
            for f in forces:
                p = np.zeros(FEXT_NPARAM)
                p[0:6] = np.asarray(f.external_forces, float).reshape(6)
                name = type(f).__name__
                if name == "ExternalForceInGlobalOnProximal":
                    k = FEXT_GLOBAL_ON_PROXIMAL
                elif name == "ExternalForceInGlobalLocalPoint":
                    k = FEXT_GLOBAL_LOCAL_POINT
                    p[6:9] = np.asarray(f.application_point_in_local, float).reshape(3)
                elif name == "ExternalForceInGlobal":
                    k = FEXT_GLOBAL
                    p[6:9] = np.asarray(f.application_point_in_global, float).reshape(3)
                elif name == "ExternalForceInLocal":
                    k = FEXT_LOCAL
                    p[6:9] = np.asarray(f.application_point_in_local, float).reshape(3)
                    p[9:18] = np.asarray(f.transformation_matrix_inv, float).reshape(9)
                else:
                    raise ValueError(name)
                seg.append(s), kind.append(k), par.append(p)
    return (
        np.array(seg, dtype=np.int32),
        np.array(kind, dtype=np.int32),
        np.array(par, dtype=np.float64).reshape(len(seg), FEXT_NPARAM),
    )


# ----------------------------------------------------------------------------- reference evaluation
def evaluate(model, Q, Qdot, fext=None, stab=(0.7, 1.3)):
    """Reference outputs for a batch of states, one system at a time (the reference has no batch axis)."""
    out = {k: [] for k in "phi_r phi_j K_j phi_h K_h bias fext Qddot lam Qddot_stab lam_stab kinetic potential".split()}
    post = {k: [] for k in "com markers dae_Qddot dae_lam".split()}
    for q, qd in zip(Q, Qdot):
        Qn, Qdn = NaturalCoordinates(q), NaturalVelocities(qd)
        out["phi_r"].append(np.asarray(model.rigid_body_constraints(Qn)).reshape(-1))
        out["phi_j"].append(np.asarray(model.joint_constraints(Qn)).reshape(-1))
        out["K_j"].append(np.asarray(model.joint_constraints_jacobian(Qn)))
        out["phi_h"].append(np.asarray(model.holonomic_constraints(Qn)).reshape(-1))
        out["K_h"].append(np.asarray(model.holonomic_constraints_jacobian(Qn)))
        out["bias"].append(np.asarray(model.holonomic_constraints_acceleration_bias(Qdn)).reshape(-1))
        out["kinetic"].append(float(np.asarray(model.kinetic_energy(Qdn)).reshape(-1)[0]))
        out["potential"].append(float(np.asarray(model.potential_energy(Qn)).reshape(-1)[0]))
        fe = model.external_force_set() if fext is None else fext
        out["fext"].append(np.asarray(fe.to_natural_external_forces(Qn)).reshape(-1))
        qdd, lam = model.forward_dynamics(Qn, Qdn, external_forces=fext)
        out["Qddot"].append(np.asarray(qdd).reshape(-1)), out["lam"].append(np.asarray(lam).reshape(-1))
        qdd, lam = model.forward_dynamics(Qn, Qdn, external_forces=fext, stabilization=dict(alpha=stab[0], beta=stab[1]))
        out["Qddot_stab"].append(np.asarray(qdd).reshape(-1)), out["lam_stab"].append(np.asarray(lam).reshape(-1))
        # post-processing / single-segment rows
        post["markers"].append(np.asarray(model.markers(Qn))[:, :, 0])  # (3, nb_markers)
        post["com"].append(np.asarray(model.center_of_mass_position(Qn))[:, :, 0])  # (3, N)
        dq, dl = [], []
        for i, seg in enumerate(model.segments_no_ground.values()):
            # natural_segment.py:574-633.  The `stabilization` branch of this function cannot run in the reference
            # (:617 subtracts a (6,) from a (6,1) in place -> ValueError), so only the plain DAE has a golden.
            a, l = seg.differential_algebraic_equation(Qn.vector(i), Qdn.vector(i))
            dq.append(np.asarray(a).reshape(-1)), dl.append(np.asarray(l).reshape(-1))
        post["dae_Qddot"].append(np.array(dq)), post["dae_lam"].append(np.array(dl))
    res = {"ref_" + k: np.array(v) for k, v in out.items()}
    res.update({"ref_" + k: np.array(v) for k, v in post.items()})
    # time_series_utils.py:6-41: norm of the residual vector per frame
    res["ref_total_rigid"] = np.sqrt(np.sum(res["ref_phi_r"] ** 2, axis=1))
    res["ref_total_joint"] = np.sqrt(np.sum(res["ref_phi_j"] ** 2, axis=1))
    res["stab_alpha_beta"] = np.array(stab)
    return res


def integrate(model, Q, Qdot, t, normalize, fext=None):
    """Reference RK4 (bionc/utils/ode_solver.py:8-53) with the closure of :107-116, per system."""
    n = model.nb_Q
    finals, mids = [], []

    def dyn(_t, y):
        qdd, _ = model.forward_dynamics(NaturalCoordinates(y[:n]), NaturalVelocities(y[n:]), external_forces=fext)
        return np.concatenate((y[n:], np.asarray(qdd).reshape(-1)), axis=0)

    for q, qd in zip(Q, Qdot):
        y = RK4(t=t, f=dyn, y0=np.concatenate([q, qd]), normalize_idx=model.normalized_coordinates if normalize else None)
        finals.append(y[:, -1]), mids.append(y[:, (len(t) - 1) // 2])
    return np.array(finals), np.array(mids)


ONLY = set(sys.argv[1:])


def write_case(name, model, Q, Qdot, fext=None, traj=None, extra=None):
    if ONLY and name not in ONLY:
        return None
    t0 = time.time()
    table = ModelTable.from_numpy(model)
    table.save(os.path.join(OUT_MODELS, name + ".npz"))
    Q, Qdot = np.atleast_2d(Q), np.atleast_2d(Qdot)
    data = dict(Q=Q, Qdot=Qdot)
    fs, fk, fp = fext_to_arrays(fext)
    data.update(fext_seg=fs, fext_kind=fk, fext_param=fp)
    data.update(evaluate(model, Q, Qdot, fext))
    data["ref_mass_matrix"] = np.asarray(model.mass_matrix)
    data["ref_gravity"] = np.asarray(model.gravity_forces()).reshape(-1)
    data["ref_max_abs_phi"] = np.abs(data["ref_phi_h"]).max()
    KQd = np.einsum("smn,sn->sm", data["ref_K_h"], Qdot)
    data["ref_max_abs_KQdot"] = np.abs(KQd).max()
    if traj is not None:
        for tag, spec in traj.items():
            ns = spec.get("systems", Q.shape[0])
            t = spec["t"]
            yf, ym = integrate(model, Q[:ns], Qdot[:ns], t, spec["normalize"], fext)
            data[f"traj_{tag}_t"] = t
            data[f"traj_{tag}_normalize"] = np.array(spec["normalize"])
            data[f"traj_{tag}_final"] = yf
            data[f"traj_{tag}_mid"] = ym
    if extra:
        data.update(extra)
    np.savez_compressed(os.path.join(OUT_CASES, name + ".npz"), **data)
    print(
        f"{name:28s} N={table.nb_segments:2d} m={table.nb_holonomic_constraints:3d} S={Q.shape[0]:2d} "
        f"|Phi|max={data['ref_max_abs_phi']:.1e} |KQd|max={data['ref_max_abs_KQdot']:.1e} "
        f"M4>0={table.is_mass_matrix_positive_definite}  ({time.time() - t0:.1f}s)"
    )
    return table


# ----------------------------------------------------------------------------- models
def box_segments():
    """The two segments of the reference's tests/test_joints.py:35-58."""
    box = NaturalSegment.with_cartesian_inertial_parameters(
        name="box", alpha=np.pi / 2, beta=np.pi / 2, gamma=np.pi / 2, length=1, mass=1,
        center_of_mass=np.array([0, 0, 0]), inertia=np.eye(3),
        inertial_transformation_matrix=TransformationMatrixType.Buv,
    )
    bbox = NaturalSegment.with_cartesian_inertial_parameters(
        name="bbox", alpha=np.pi / 1.9, beta=np.pi / 2.3, gamma=np.pi / 2.1, length=1.5, mass=1.1,
        center_of_mass=np.array([0.1, 0.11, 0.111]), inertia=np.diag([1.1, 1.2, 1.3]),
        inertial_transformation_matrix=TransformationMatrixType.Buv,
    )
    return box, bbox


TJ_Q1 = np.array([1, 2, 3.05, 1.1, 1, 3.1, 1.2, 2, 4.1, 1.3, 2, 5.1])  # tests/test_joints.py:321-326
TJ_Q2 = np.array([1.4, 2.1, 3.2, 1.5, 1.1, 3.2, 1.6, 2.2, 4.2, 1.7, 2, 5.3])  # tests/test_joints.py:327-332


def test_joints_models():
    """One two-segment model per joint type of the reference's tests/test_joints.py, evaluated at
    the test's (non-consistent) Q1, Q2; the velocities reuse Q1, Q2 as the test does (:371)."""
    specs = {
        "tj_revolute": dict(joint_type=JointType.REVOLUTE, parent="box", child="bbox",
                            parent_axis=(NaturalAxis.U, NaturalAxis.V), child_axis=(NaturalAxis.V, NaturalAxis.W),
                            theta=(np.pi / 3, 3 * np.pi / 4)),
        "tj_universal": dict(joint_type=JointType.UNIVERSAL, parent="box", child="bbox",
                             parent_axis=NaturalAxis.U, child_axis=NaturalAxis.W, theta=0.4),
        "tj_spherical": dict(joint_type=JointType.SPHERICAL, parent="box", child="bbox"),
        "tj_spherical_custom": dict(joint_type=JointType.SPHERICAL, parent="box", child="bbox",
                                    parent_point="P1", child_point="P2"),
        "tj_constant_length": dict(joint_type=JointType.CONSTANT_LENGTH, parent="box", child="bbox",
                                   parent_point="P1", child_point="P2", length=1.5),
        "tj_sphere_on_plane": dict(joint_type=JointType.SPHERE_ON_PLANE, parent="box", child="bbox",
                                   sphere_radius=0.02, sphere_center="P1", plane_point="P2", plane_normal="PLANE_NORMAL"),
        "tj_ground_revolute": dict(joint_type=JointType.GROUND_REVOLUTE, parent="GROUND", child="box",
                                   parent_axis=[CartesianAxis.X, CartesianAxis.X],
                                   child_axis=[NaturalAxis.V, NaturalAxis.W], theta=[np.pi / 2, np.pi / 2]),
        "tj_ground_spherical": dict(joint_type=JointType.GROUND_SPHERICAL, parent="GROUND", child="box",
                                    ground_application_point=np.array([0.1, 0.2, 0.3])),
        "tj_ground_weld": dict(joint_type=JointType.GROUND_WELD, parent="GROUND", child="bbox",
                               rp_child_ref=np.array([0.1, 0.2, 0.3]), rd_child_ref=np.array([0.2, 0.04, 0.05])),
        "tj_ground_universal": dict(joint_type=JointType.GROUND_UNIVERSAL, parent="GROUND", child="bbox",
                                    parent_axis=CartesianAxis.X, child_axis=NaturalAxis.V, theta=np.pi / 2),
    }
    for name, spec in specs.items():
        box, bbox = box_segments()
        box.add_natural_marker_from_segment_coordinates(name="P1", location=[0.1, 0.2, 0.3], is_anatomical=True)
        bbox.add_natural_marker_from_segment_coordinates(name="P2", location=[0.2, 0.04, 0.05], is_anatomical=True)
        bbox.add_natural_vector_from_segment_coordinates(name="PLANE_NORMAL", direction=np.array([0.2, 0.04, 0.05]), normalize=True)
        model = BiomechanicalModel()
        model["box"] = box
        model["bbox"] = bbox
        model._add_joint(dict(name=name, **spec))
        Q = np.concatenate([TJ_Q1, TJ_Q2])
        # the KKT at this non-consistent state is fine for every joint type (probe); velocities = Q
        write_case(name, model, Q, Q.copy())


def lower_limb_model():
    """Config 3: geometry of examples/models/lower_limb.nc, joints of
    examples/model_creation/right_side_lower_limb.py:226-269, inertia assigned here."""
    src = BiomechanicalModel.load(os.path.join(REF, "examples/models/lower_limb.nc"))
    inertia = {
        "PELVIS": (10.65, (0, -0.05, 0), (0.10, 0.11, 0.09)),
        "THIGH": (9.2, (0, -0.17, 0), (0.13, 0.05, 0.13)),
        "SHANK": (3.6, (0, -0.10, 0), (0.040, 0.020, 0.040)),
        "FOOT": (0.9, (0.02, -0.03, 0), (0.0040, 0.0045, 0.0045)),
    }
    for name, (m, c, I) in inertia.items():
        src.segments[name].set_inertial_parameters(
            mass=m, center_of_mass=np.array(c, float), inertia_matrix=np.diag(I),
            transformation_matrix=TransformationMatrixType.Buv,
        )
    src._update_mass_matrix()
    return src


FULL_BODY_SEGMENTS = [
    # name, parent, (L, alpha, beta, gamma), (mass, com_scs, diag inertia), parent attach point in parent SCS (None = distal)
    ("PELVIS", None, (0.1770, 1.9763, 1.6390, 2.4511), (10.65, (0, -0.05, 0), (0.10, 0.11, 0.09)), None),
    ("R_THIGH", "PELVIS", (0.3965, 1.2764, np.pi / 2, np.pi / 2), (9.2, (0, -0.17, 0), (0.13, 0.05, 0.13)), (0.0, -0.08, 0.09)),
    ("R_SHANK", "R_THIGH", (0.3531, 1.8364, np.pi / 2, np.pi / 2), (3.6, (0, -0.10, 0), (0.040, 0.020, 0.040)), None),
    ("R_FOOT", "R_SHANK", (0.1389, 1.7427, 1.4427, 2.8024), (0.9, (0.02, -0.03, 0), (0.0040, 0.0045, 0.0045)), None),
    ("L_THIGH", "PELVIS", (0.3965, 1.2764, np.pi / 2, np.pi / 2), (9.2, (0, -0.17, 0), (0.13, 0.05, 0.13)), (0.0, -0.08, -0.09)),
    ("L_SHANK", "L_THIGH", (0.3531, 1.8364, np.pi / 2, np.pi / 2), (3.6, (0, -0.10, 0), (0.040, 0.020, 0.040)), None),
    ("L_FOOT", "L_SHANK", (0.1389, 1.7427, 1.4427, 2.8024), (0.9, (0.02, -0.03, 0), (0.0040, 0.0045, 0.0045)), None),
    ("THORAX", "PELVIS", (0.45, np.pi / 2, np.pi / 2, np.pi / 2), (25.0, (0, 0.20, 0), (0.60, 0.30, 0.60)), (0.0, 0.10, 0.0)),
    ("R_ARM", "THORAX", (0.30, np.pi / 2, np.pi / 2, np.pi / 2), (2.1, (0, -0.08, 0), (0.060, 0.012, 0.060)), (0.0, -0.9, 0.20)),
    ("R_FOREARM", "R_ARM", (0.27, np.pi / 2, np.pi / 2, np.pi / 2), (1.2, (0, -0.04, 0), (0.030, 0.005, 0.030)), None),
]


def full_body_model():
    """Config 5 (synthetic; no such model exists in the reference): 10 segments, a free pelvis and
    9 spherical joints, geometry of lower_limb.nc for both legs plus a thorax and a right arm."""
    model = BiomechanicalModel()
    for name, parent, (L, al, be, ga), (m, c, I), attach in FULL_BODY_SEGMENTS:
        model[name] = NaturalSegment.with_cartesian_inertial_parameters(
            name=name, alpha=al, beta=be, gamma=ga, length=L, mass=m,
            center_of_mass=np.array(c, float), inertia=np.diag(I),
            inertial_transformation_matrix=TransformationMatrixType.Buv,
        )
    for name, parent, _, _, attach in FULL_BODY_SEGMENTS:
        if parent is None:
            continue
        kw = {}
        if attach is not None:
            model[parent].add_natural_marker_from_segment_coordinates(
                name=f"attach_{name}", location=np.array(attach, float), is_anatomical=True
            )
            kw["parent_point"] = f"attach_{name}"
        model._add_joint(dict(name=f"joint_{name}", joint_type=JointType.SPHERICAL, parent=parent, child=name, **kw))
    return model


def main():
    os.makedirs(OUT_MODELS, exist_ok=True)
    os.makedirs(OUT_CASES, exist_ok=True)
    fd = "examples/forward_dynamics/"

    # ---- joint zoo of the reference's tests/test_joints.py
    test_joints_models()

    # ---- config 1: root pendulum.nmod (z_revolute variant, indefinite G; examples/forward_dynamics/pendulum.py:195-201)
    model = BiomechanicalModel.load(os.path.join(REF, "pendulum.nmod"))
    Q0 = np.array([0, 0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0], float)
    # swing it about the hinge (global z for this variant: axes X,X against U,V -> w stays along x?)  keep
    # the example's own initial state plus rotated copies about the free axis found numerically below
    t1000 = np.linspace(0, 10, 1001)
    write_case(
        "pendulum_root", model, Q0, np.zeros(12),
        traj=dict(plain=dict(t=t1000, normalize=False), norm=dict(t=t1000, normalize=True)),
    )

    # ---- the three pendulum.py modes, 500 steps, no renormalisation (tests/test_ultimate_examples.py:74-146)
    mod = load_example(fd + "pendulum.py")
    for mode in ("x_revolute", "y_revolute", "z_revolute"):
        model, all_states, _ = mod.main(mode=mode, show_results=False)
        t = np.linspace(0, 10, 501)
        write_case(
            f"pendulum_{mode}", model, all_states[:12, 0], all_states[12:, 0],
            traj=dict(example=dict(t=t, normalize=False)),
            extra=dict(example_final=all_states[:, -1]),
        )

    # ---- config 2: pendulum_with_force.nmod, 3 modes, 600 steps with renormalisation (tests/test_ultimate_examples.py:149-249)
    mod = load_example(fd + "pendulum_with_force.py")
    for mode in ("moment_equilibrium", "force_equilibrium", "no_equilibrium"):
        model, all_states, _ = mod.main(mode=mode)
        fext = ExternalForceSet.empty_from_nb_segment(1)
        if mode == "moment_equilibrium":
            fext.add_in_global_local_point(external_force=np.array([-9.81 * 0.5, 0, 0, 0, 0, 0]), segment_index=0, point_in_local=np.array([0, 0, 0]))
        elif mode == "force_equilibrium":
            fext.add_in_global_local_point(external_force=np.array([0, 0, 0, 0, 0, 9.81]), segment_index=0, point_in_local=np.array([0, 0.5, 0]))
        else:
            fext.add_in_global_local_point(external_force=np.array([0, 0, 9.81, 0, 0, 0]), segment_index=0, point_in_local=np.array([0, 0.25, 0]))
        t = np.linspace(0, 10, 601)
        write_case(
            f"pendulum_force_{mode}", model, all_states[:12, 0], all_states[12:, 0], fext=fext,
            traj=dict(example=dict(t=t, normalize=True)),
            extra=dict(example_final=all_states[:, -1]),
        )
    # config-2 batch workload: random hinge states, the no_equilibrium force, h = 1/60, 1000 steps
    model = BiomechanicalModel.load(os.path.join(REF, "pendulum_with_force.nmod"))
    fext = ExternalForceSet.empty_from_nb_segment(1)
    fext.add_in_global_local_point(external_force=np.array([0, 0, 9.81, 0, 0, 0]), segment_index=0, point_in_local=np.array([0, 0.25, 0]))
    Q, Qd = st.hinge_pendulum_states(8, seed=0)
    write_case(
        "pendulum_force_batch", model, Q, Qd, fext=fext,
        traj=dict(cfg=dict(t=np.arange(1001) / 60.0, normalize=True, systems=3)),
    )
    # all four external-force kinds at once (tests/test_external_force.py covers them one by one)
    fext = ExternalForceSet.empty_from_nb_segment(1)
    fext.add_in_global_local_point(external_force=np.array([0.3, -0.2, 0.5, 1.0, 2.0, -3.0]), segment_index=0, point_in_local=np.array([0.1, 0.25, -0.2]))
    fext.add_in_global(external_force=np.array([0.1, 0.4, -0.3, -1.5, 0.5, 2.5]), segment_index=0, point_in_global=np.array([0.3, -0.4, 0.2]))
    fext.add_in_local(external_force=np.array([-0.2, 0.1, 0.6, 0.7, -0.9, 1.1]), segment_index=0, point_in_local=np.array([0.05, -0.5, 0.1]),
                      transformation_matrix=model.segments["pendulum"].compute_transformation_matrix(TransformationMatrixType.Buv))
    fext.external_forces[0].append(bionc.bionc_numpy.external_force_global_on_proximal.ExternalForceInGlobalOnProximal(np.array([0.2, 0.3, -0.1, 0.4, -0.6, 0.8])))
    write_case("pendulum_force_allkinds", model, Q[:4], Qd[:4], fext=fext,
               traj=dict(short=dict(t=np.arange(101) / 100.0, normalize=True, systems=2)))

    # ---- ground universal pendulum and double universal pendulum (saved models of the examples)
    model = BiomechanicalModel.load(os.path.join(REF, fd + "pendulum.nmod"))
    rng = np.random.default_rng(5)
    Qs, Qds = [], []
    for _ in range(4):
        # ground universal (X vs V, theta = pi/2): v orthogonal to x, free rotation about x and about v
        R = st.rotation_from_rotvec(np.array([rng.uniform(-1, 1), 0, 0]))
        spin = rng.uniform(-1, 1)
        frame = np.eye(3)
        Rv = st.rotation_from_rotvec(spin * (R @ frame)[:, 1])
        om = np.array([rng.normal(), 0, 0]) + rng.normal() * (R @ frame)[:, 1]
        q, qd = st._place((Rv @ R)[None], frame, np.zeros((1, 3)), om[None], np.zeros((1, 3)))
        Qs.append(q[0]), Qds.append(qd[0])
    write_case("pendulum_universal", model, np.array(Qs), np.array(Qds),
               traj=dict(short=dict(t=np.linspace(0, 2, 201), normalize=False, systems=2)))

    mod = load_example(fd + "double_pendulum_with_force.py")
    for mode in ("force_equilibrium", "no_equilibrium"):
        model, all_states, _ = mod.main(mode=mode)
        fext = ExternalForceSet.empty_from_nb_segment(nb_segment=2)
        pt = [0, 0.5, 0] if mode == "force_equilibrium" else [0, 0.25, 0]
        for s in range(2):
            fext.add_in_global_local_point(segment_index=s, external_force=np.array([0, 0, 0, 0, 0, 9.81]), point_in_local=np.array(pt))
        write_case(
            f"double_pendulum_force_{mode}", model, all_states[:24, 0], all_states[24:, 0], fext=fext,
            traj=dict(example=dict(t=np.linspace(0, 10, 601), normalize=True)),
            extra=dict(example_final=all_states[:, -1]),
        )

    # ---- 4-link hinge pendulum (tests/test_forward_dynamics.py:251-501 uses this builder)
    mod = load_example(fd + "n_link_pendulum.py")
    model = mod.build_n_link_pendulum(nb_segments=4)
    table = ModelTable.from_numpy(model)
    Q, Qd = st.planar_chain_states(table, 6, seed=1)
    # plus the non-consistent linspace state of the reference test (:268-269)
    Ql = np.linspace(0, 0.24, 48)
    Qdl = np.linspace(0, 0.02, 48)
    write_case("n_link_4", model, Q, Qd, traj=dict(short=dict(t=np.linspace(0, 1, 201), normalize=True, systems=2)),
               extra=dict(Q_linspace=Ql, Qdot_linspace=Qdl,
                          ref_Qddot_linspace=np.asarray(model.forward_dynamics(NaturalCoordinates(Ql), NaturalVelocities(Qdl))[0]).reshape(-1),
                          ref_K_linspace=np.asarray(model.holonomic_constraints_jacobian(NaturalCoordinates(Ql))),
                          ref_bias_linspace=np.asarray(model.holonomic_constraints_acceleration_bias(NaturalVelocities(Qdl))).reshape(-1)))

    # ---- config 3: lower limb
    model = lower_limb_model()
    table = ModelTable.from_numpy(model)
    Q, Qd = st.tree_states(table, 8, seed=0)
    write_case("lower_limb", model, Q, Qd,
               traj=dict(cfg=dict(t=np.linspace(0, 10, 1001), normalize=True, systems=2),
                         fine=dict(t=np.linspace(0, 1, 1001), normalize=True, systems=2)))

    # ---- config 4: knee parallel mechanism (Feikes)
    mod = load_example("examples/knee_parallel_mechanism/knee_feikes.py")
    model = mod.create_knee_model()
    Qk = np.array([0, 0, 1, 0, 0.0, 0, 0, 0.4, 0, -1, 0, 0, 0, 0, 1, 0, 0.4, 0, 0, 0.8, 0, -1, 0, 0], float)
    Q, Qd = st.rigidly_moved_states(Qk, 6, seed=0, angle_max=0.3, omega_sigma=0.5)
    Q[0], Qd[0] = Qk, 0.0  # the example's own state (simulation_feikes.py:11-14), at rest
    write_case("knee_feikes", model, Q, Qd,
               traj=dict(cfg=dict(t=np.arange(1001) * 5e-5, normalize=True, systems=3)))

    # ---- config 5: synthetic full body, nominal and per-system perturbed inertial parameters
    model = full_body_model()
    table = ModelTable.from_numpy(model)
    Q, Qd = st.tree_states(table, 4, seed=2)
    write_case("full_body", model, Q, Qd, traj=dict(cfg=dict(t=np.arange(101) * 0.01, normalize=True, systems=2)))
    mass, com, J = st.perturbed_inertial_parameters(table, 4, seed=3)
    qdds, lams, finals = [], [], []
    for s in range(4):
        ms = full_body_model()
        for i, seg in enumerate(ms.segments_no_ground.values()):
            seg.set_natural_inertial_parameters(mass=mass[s, i], natural_center_of_mass=com[s, i], natural_pseudo_inertia=J[s, i])
        ms._update_mass_matrix()
        qdd, lam = ms.forward_dynamics(NaturalCoordinates(Q[s]), NaturalVelocities(Qd[s]))
        qdds.append(np.asarray(qdd).reshape(-1)), lams.append(np.asarray(lam).reshape(-1))
        if s < 2:
            yf, _ = integrate(ms, Q[s : s + 1], Qd[s : s + 1], np.arange(101) * 0.01, True)
            finals.append(yf[0])
    np.savez_compressed(
        os.path.join(OUT_CASES, "full_body_perturbed.npz"),
        Q=Q, Qdot=Qd, seg_mass=mass, seg_com=com, seg_J=J,
        ref_Qddot=np.array(qdds), ref_lam=np.array(lams),
        traj_cfg_t=np.arange(101) * 0.01, traj_cfg_final=np.array(finals),
    )
    print("full_body_perturbed          done")

    # ---- single free box (tests/test_forward_dynamics.py:22-241; drop_the_box.py), as a 1-segment model
    model = BiomechanicalModel()
    model["box"] = NaturalSegment.with_cartesian_inertial_parameters(
        name="box", alpha=np.pi / 2, beta=np.pi / 2, gamma=np.pi / 2, length=1, mass=1,
        center_of_mass=np.array([0, 0, 0]), inertia=np.eye(3),
        inertial_transformation_matrix=TransformationMatrixType.Buv,
    )
    Qb = np.array([1, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 1], float)
    write_case("free_box", model, Qb, np.zeros(12), traj=dict(example=dict(t=np.linspace(0, 2, 101), normalize=False)))


if __name__ == "__main__":
    main()
