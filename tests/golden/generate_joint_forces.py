"""Golden vectors for joint generalized forces, from the UNMODIFIED reference.

The reference's ``forward_dynamics`` ignores joint generalized forces (call site commented out,
``bionc_numpy/biomechanical_model.py:391-413``) and ``JointGeneralizedForces.to_natural_joint_forces``
(``generalized_force.py:98-129``) cannot run: it calls ``to_natural_force`` and ``transport_to``, which no class defines
(AttributeError, checked below).  The vectors are therefore composed from the reference methods that DO exist and that
the docstrings of :98-129 describe:

    child  : ExternalForceInGlobalLocalPoint(tau, F, point = 0).to_generalized_natural_forces(Q_child)
    parent : - ExternalForceInGlobalOnProximal(tau, F).transport_to_another_segment(Q_child, Q_parent)
                 .to_generalized_natural_forces(Q_parent)

and the accelerations from the reference's own augmented matrix with ``forces = gravity + f_ext + natural_joint_forces``
(the three lines :404-413 with the commented term switched on).

    cd /tmp && python /root/repo/tests/golden/generate_joint_forces.py
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "baseline")]
import reference_models as rm  # noqa: E402

rm.import_reference()
from bionc.bionc_numpy import NaturalCoordinates, NaturalVelocities  # noqa: E402
from bionc.bionc_numpy.external_force_global_local_point import ExternalForceInGlobalLocalPoint  # noqa: E402
from bionc.bionc_numpy.external_force_global_on_proximal import ExternalForceInGlobalOnProximal  # noqa: E402
from bionc.bionc_numpy.generalized_force import JointGeneralizedForces  # noqa: E402

from bionc_b200.bionc_cuda.model_table import ModelTable  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402


def main():
    model = rm.lower_limb_model()
    table = ModelTable.from_numpy(model)
    Q, Qd = st.tree_states(table, 6, seed=4)
    joints = list(model.joints.values())
    rng = np.random.default_rng(9)
    jf = []  # (joint index, [tau, F])
    for j, joint in enumerate(joints):
        if joint.parent is None and rng.random() < 0.5:
            continue
        jf.append((j, rng.normal(0.0, 3.0, 6)))
        if rng.random() < 0.4:
            jf.append((j, rng.normal(0.0, 1.0, 6)))
    # the reference's own method cannot run
    try:
        JointGeneralizedForces(external_forces=jf[0][1], application_point_in_local=np.zeros(3)).to_natural_joint_forces(
            parent_segment_index=0, child_segment_index=1, Q=NaturalCoordinates(Q[0]))
        raise SystemExit("the reference's to_natural_joint_forces runs now: regenerate this golden from it")
    except AttributeError as e:
        print("reference JointGeneralizedForces.to_natural_joint_forces:", e)
    nat, qdd, lam = [], [], []
    for s in range(Q.shape[0]):
        Qn, Qdn = NaturalCoordinates(Q[s]), NaturalVelocities(Qd[s])
        f = np.zeros((model.nb_Q, 1))
        for j, val in jf:
            joint = joints[j]
            c = joint.child.index
            f[12 * c : 12 * c + 12] += np.asarray(ExternalForceInGlobalLocalPoint(external_forces=val, application_point_in_local=np.zeros(3))
                                                  .to_generalized_natural_forces(Qn.vector(c))).reshape(12, 1)
            if joint.parent is not None:
                p = joint.parent.index
                moved = ExternalForceInGlobalOnProximal(external_forces=val).transport_to_another_segment(Qfrom=Qn.vector(c), Qto=Qn.vector(p))
                f[12 * p : 12 * p + 12] -= np.asarray(moved.to_generalized_natural_forces(Qn.vector(p))).reshape(12, 1)
        nat.append(f[:, 0].copy())
        # bionc_numpy/biomechanical_model.py:399-426 with "+ natural_joint_forces" switched on
        A = model.augmented_mass_matrix(Qn)
        forces = model.gravity_forces() + f
        bias = -model.holonomic_constraints_acceleration_bias(Qdn)
        x = np.linalg.solve(A, np.concatenate((forces, bias), axis=0))
        qdd.append(x[: model.nb_Q, 0]), lam.append(x[model.nb_Q :, 0])
    out = os.path.join(REPO, "tests", "golden", "cases_extra")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(
        os.path.join(out, "joint_forces_lower_limb.npz"), Q=Q, Qdot=Qd,
        joint_index=np.array([j for j, _ in jf], dtype=np.int32), joint_force=np.array([v for _, v in jf]),
        all_joint_parent=table.all_joint_parent, all_joint_child=table.all_joint_child,
        ref_natural_joint_forces=np.array(nat), ref_Qddot=np.array(qdd), ref_lam=np.array(lam),
    )
    print("joints:", [(j.name, None if j.parent is None else j.parent.index, j.child.index) for j in joints])
    print("wrote joint_forces_lower_limb.npz with", len(jf), "joint forces")


if __name__ == "__main__":
    main()
