"""The LIVE reference (unmodified Ipuch/bioNC, installed by ``baseline/install_reference.py`` into ``baseline/_ref``)
and the models of BASELINE.json's configurations built with its own API.

Test / benchmark infrastructure only: ``tests/`` (live-reference parity), ``bench.py``'s reference arm and
``cpu_baseline`` leg import this; nothing under ``bionc_b200/`` does.  ``$BIONC_REFERENCE`` may point at another
checkout of the reference (the golden generator uses ``/root/reference``); the default is ``baseline/_ref``, the only
copy that exists on the GPU box.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_INSTALL = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_INSTALL, "bionc", "__init__.py")) or bool(os.environ.get("BIONC_REFERENCE"))


def reference_root() -> str:
    """Directory that holds the ``bionc`` package."""
    return os.environ.get("BIONC_REFERENCE") or REF_INSTALL


def data_path(rel: str) -> str:
    """A data file of the reference (``.nmod`` pickles, example modules): ``baseline/_ref/data/<rel>`` or, for a
    source checkout, ``<root>/<rel>``."""
    root = reference_root()
    cand = os.path.join(root, "data", rel)
    return cand if os.path.exists(cand) else os.path.join(root, rel)


def import_reference():
    """``import bionc`` (the unmodified reference) through the import shims for the absent third-party packages."""
    stubs = os.path.join(REPO, "oracle", "stubs")
    for p in (reference_root(), stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    import bionc  # noqa: F401

    return bionc


def load_example(relpath: str):
    path = data_path(relpath)
    d = os.path.dirname(path)
    if d not in sys.path:
        sys.path.insert(0, d)
    spec = importlib.util.spec_from_file_location(os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LOWER_LIMB_INERTIA = {  # (mass, centre of mass in the SCS, diagonal inertia): SURVEY 8(d), validated eig(M4) > 0
    "PELVIS": (10.65, (0, -0.05, 0), (0.10, 0.11, 0.09)),
    "THIGH": (9.2, (0, -0.17, 0), (0.13, 0.05, 0.13)),
    "SHANK": (3.6, (0, -0.10, 0), (0.040, 0.020, 0.040)),
    "FOOT": (0.9, (0.02, -0.03, 0), (0.0040, 0.0045, 0.0045)),
}

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

KNEE_Q0 = np.array([0, 0, 1, 0, 0.0, 0, 0, 0.4, 0, -1, 0, 0, 0, 0, 1, 0, 0.4, 0, 0, 0.8, 0, -1, 0, 0], float)  # simulation_feikes.py:11-14


def lower_limb_model():
    """Config 3: geometry of examples/models/lower_limb.nc, joints of
    examples/model_creation/right_side_lower_limb.py:226-269, inertia assigned here."""
    import_reference()
    from bionc import TransformationMatrixType
    from bionc.bionc_numpy import BiomechanicalModel

    src = BiomechanicalModel.load(data_path("examples/models/lower_limb.nc"))
    for name, (m, c, I) in LOWER_LIMB_INERTIA.items():
        src.segments[name].set_inertial_parameters(
            mass=m, center_of_mass=np.array(c, float), inertia_matrix=np.diag(I),
            transformation_matrix=TransformationMatrixType.Buv,
        )
    src._update_mass_matrix()
    return src


def full_body_model():
    """Config 5 (synthetic; no such model exists in the reference): 10 segments, a free pelvis and
    9 spherical joints, geometry of lower_limb.nc for both legs plus a thorax and a right arm."""
    import_reference()
    from bionc import JointType, TransformationMatrixType
    from bionc.bionc_numpy import BiomechanicalModel, NaturalSegment

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


def knee_model():
    """Config 4: ``create_knee_model()`` of examples/knee_parallel_mechanism/knee_feikes.py:11-251."""
    import_reference()
    return load_example("examples/knee_parallel_mechanism/knee_feikes.py").create_knee_model()


def pendulum_model():
    """Config 1: the saved ``pendulum.nmod`` at the reference root (ground hinge, indefinite G)."""
    import_reference()
    from bionc.bionc_numpy import BiomechanicalModel

    return BiomechanicalModel.load(data_path("pendulum.nmod"))


def pendulum_with_force_model():
    """Config 2: ``pendulum_with_force.nmod`` and the ``no_equilibrium`` force set of
    examples/forward_dynamics/pendulum_with_force.py:188-194 -> (model, ExternalForceSet)."""
    import_reference()
    from bionc.bionc_numpy import BiomechanicalModel, ExternalForceSet

    model = BiomechanicalModel.load(data_path("pendulum_with_force.nmod"))
    fext = ExternalForceSet.empty_from_nb_segment(1)
    fext.add_in_global_local_point(external_force=np.array([0, 0, 9.81, 0, 0, 0]), segment_index=0, point_in_local=np.array([0, 0.25, 0]))
    return model, fext


# name -> (builder, step size of the configuration, state generator kind)
CONFIGS = {
    "pendulum": dict(build=lambda: (pendulum_model(), None), h=0.01, states="pendulum_root"),
    "pendulum_force": dict(build=pendulum_with_force_model, h=1.0 / 60.0, states="hinge"),
    "lower_limb": dict(build=lambda: (lower_limb_model(), None), h=0.01, states="tree"),
    "knee": dict(build=lambda: (knee_model(), None), h=5e-5, states="knee"),
    "full_body": dict(build=lambda: (full_body_model(), None), h=0.001, states="tree"),
}


def config_states(name: str, table, count: int, seed: int):
    """Seeded, constraint-consistent initial states of a configuration (SURVEY 8(d) generators, shared with bench.py)."""
    from bionc_b200.utils import states as st

    kind = CONFIGS[name]["states"]
    if kind == "tree":
        return st.tree_states(table, count, seed=seed)
    if kind == "hinge":
        return st.hinge_pendulum_states(count, seed=seed)
    if kind == "knee":
        return st.rigidly_moved_states(KNEE_Q0, count, seed=seed, angle_max=0.3, omega_sigma=0.0)
    # pendulum.nmod: the example's own initial state (examples/forward_dynamics/pendulum.py:195-201) swung about the
    # hinge axis (global x for this variant's axes) with a random rate
    rng = np.random.default_rng(seed)
    q0 = np.array([0, 0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0], float)
    Q, Qd = np.tile(q0, (count, 1)), np.zeros((count, 12))
    return Q, Qd


def reference_rk4(model, Q0, Qdot0, t, normalize=True, external_forces=None, stabilization=None):
    """The reference's own integrator on one system: ``RK4`` (utils/ode_solver.py:8-53) around the closure of
    ``forward_integration`` (:107-116) -> final (Q, Qdot)."""
    from bionc.bionc_numpy import NaturalCoordinates, NaturalVelocities
    from bionc.utils.ode_solver import RK4

    n = model.nb_Q

    def dynamics(_t, states):
        qddot, _lam = model.forward_dynamics(NaturalCoordinates(states[:n]), NaturalVelocities(states[n:]),
                                             external_forces=external_forces, stabilization=stabilization)
        return np.concatenate((states[n:], qddot.to_array()), axis=0)

    y = RK4(t, dynamics, np.concatenate([np.asarray(Q0, float).reshape(-1), np.asarray(Qdot0, float).reshape(-1)]),
            normalize_idx=model.normalized_coordinates if normalize else None)
    return y[:n, -1], y[n:, -1]
