"""Joint generalized forces for the CUDA backend: the construction API of the reference's
``JointGeneralizedForces`` / ``JointGeneralizedForcesList`` (``bionc/bionc_numpy/generalized_force.py:11-129, :159-311``)
lowered to the kernels' force table.

A joint force is the wrench ``[torque(3), force(3)]`` (global frame) the joint applies to its CHILD segment at the child's
proximal point; the joint's PARENT (if it is not the ground) receives the opposite wrench transported to its own
proximal point (``to_natural_joint_forces``, :98-129).  The reference's ``forward_dynamics`` ignores joint generalized
forces -- the call site is commented out (``bionc_numpy/biomechanical_model.py:391-413``) and
``JointGeneralizedForces.to_natural_joint_forces`` itself cannot run (it calls ``to_natural_force`` / ``transport_to``,
which no class defines) -- so the backend ignores them as well unless ``apply_joint_generalized_forces=True`` is
passed.  The semantics then are the documented ones, composed from the reference methods that do exist:
child ``ExternalForceInGlobalLocalPoint(...).to_generalized_natural_forces(Q_child)``, parent minus
``ExternalForceInGlobalOnProximal.transport_to_another_segment(Q_child, Q_parent)``
(``external_force_global_on_proximal.py:117-147``, lever arm and sign as coded there); the golden of
``tests/golden/generate_joint_forces.py`` is produced that way.
"""

from __future__ import annotations

import numpy as np

from .model_table import FEXT_GLOBAL_LOCAL_POINT, FEXT_JOINT_REACTION, FEXT_NPARAM


class JointGeneralizedForces:
    """``[torque, force]`` of one joint, applied at the child's proximal point (generalized_force.py:11-42)."""

    def __init__(self, external_forces, application_point_in_local=None):
        self.external_forces = np.asarray(external_forces, dtype=np.float64).reshape(6)
        pt = np.zeros(3) if application_point_in_local is None else np.asarray(application_point_in_local, dtype=np.float64).reshape(3)
        if np.any(pt != 0.0):
            raise NotImplementedError("joint generalized forces act at the child's proximal point (the reference always builds them "
                                      "with application_point_in_local = 0, generalized_force.py:93-96)")
        self.application_point_in_local = pt

    @property
    def torque(self):
        return self.external_forces[0:3]

    @property
    def force(self):
        return self.external_forces[3:6]


class JointGeneralizedForcesList:
    """One list of forces per joint, in the order of the model's joints (generalized_force.py:159-230)."""

    def __init__(self, joint_generalized_forces: list[list] = None):
        if joint_generalized_forces is None:
            raise ValueError(
                "joint_generalized_forces must be a list of JointGeneralizedForces, or use the classmethod"
                "NaturalExternalForceSet.empty_from_nb_joint(nb_joint)"
            )
        self.joint_generalized_forces = joint_generalized_forces

    @property
    def nb_joint(self) -> int:
        return len(self.joint_generalized_forces)

    @classmethod
    def empty_from_nb_joint(cls, nb_joint: int):
        return cls(joint_generalized_forces=[[] for _ in range(nb_joint)])

    def joint_generalized_force(self, joint_index: int) -> list:
        return self.joint_generalized_forces[joint_index]

    def add_generalized_force(self, joint_index: int, joint_generalized_force):
        if not isinstance(joint_generalized_force, JointGeneralizedForces):  # accept the reference's object or a 6-vector
            ef = getattr(joint_generalized_force, "external_forces", joint_generalized_force)
            pt = getattr(joint_generalized_force, "application_point_in_local", None)
            joint_generalized_force = JointGeneralizedForces(ef, pt)
        self.joint_generalized_forces[joint_index].append(joint_generalized_force)

    @classmethod
    def from_reference(cls, jgf) -> "JointGeneralizedForcesList":
        out = cls.empty_from_nb_joint(len(jgf.joint_generalized_forces))
        for j, forces in enumerate(jgf.joint_generalized_forces):
            for f in forces:
                out.add_generalized_force(j, f)
        return out

    def to_arrays(self, joint_parent, joint_child):
        """(segment, kind, param) rows of the kernels' force table: per force one entry on the child and, unless the
        parent is the ground, its reaction on the parent (the child's index travels in param[18])."""
        if len(joint_parent) != self.nb_joint:
            # same check as generalized_force.py:255-262
            raise ValueError(
                "The number of joint in the model and the number of segment in the joint forces must be the same."
                f"Got {self.nb_joint} joint forces and {len(joint_parent)} joints in the model."
            )
        seg, kind, par = [], [], []
        for j, forces in enumerate(self.joint_generalized_forces):
            p, c = int(joint_parent[j]), int(joint_child[j])
            for f in forces:
                q = np.zeros(FEXT_NPARAM)
                q[0:6] = f.external_forces
                seg.append(c), kind.append(FEXT_GLOBAL_LOCAL_POINT), par.append(q)
                if p >= 0:
                    r = q.copy()
                    r[18] = c
                    seg.append(p), kind.append(FEXT_JOINT_REACTION), par.append(r)
        return (np.ascontiguousarray(seg, dtype=np.int32), np.ascontiguousarray(kind, dtype=np.int32),
                np.ascontiguousarray(np.array(par, dtype=np.float64).reshape(len(seg), FEXT_NPARAM)))
