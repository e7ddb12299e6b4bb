"""``ExternalForceSet`` for the CUDA backend: same construction API as the reference
(``bionc/bionc_numpy/external_force.py:13-180``), lowered to flat arrays the kernels consume.

A force is the 6-vector ``[torque(3), force(3)]`` (torque first, ``external_force.py:84-85``),
expressed in the global frame (or in the segment frame for ``add_in_local``), with an application
point.  The same set applies to every system of a batch.
"""

from __future__ import annotations

import numpy as np

from .model_table import FEXT_GLOBAL, FEXT_GLOBAL_LOCAL_POINT, FEXT_GLOBAL_ON_PROXIMAL, FEXT_LOCAL, FEXT_NPARAM


class ExternalForceSet:
    def __init__(self, external_forces: list[list] = None):
        if external_forces is None:
            raise ValueError(
                "f_ext must be a list of ExternalForces, or use the classmethod"
                "NaturalExternalForceSet.empty_from_nb_segment(nb_segment)"
            )
        # list (per segment) of lists of (kind, param[FEXT_NPARAM])
        self.external_forces = external_forces

    @property
    def nb_segments(self) -> int:
        return len(self.external_forces)

    @classmethod
    def empty_from_nb_segment(cls, nb_segment: int):
        return cls(external_forces=[[] for _ in range(nb_segment)])

    def segment_external_forces(self, segment_index: int) -> list:
        return self.external_forces[segment_index]

    @staticmethod
    def _param(external_force, point=None, tinv=None):
        p = np.zeros(FEXT_NPARAM)
        p[0:6] = np.asarray(external_force, dtype=np.float64).reshape(6)
        if point is not None:
            p[6:9] = np.asarray(point, dtype=np.float64).reshape(3)
        if tinv is not None:
            p[9:18] = np.asarray(tinv, dtype=np.float64).reshape(9)
        return p

    def add_in_global(self, segment_index: int, external_force, point_in_global=None):
        """external_force_global.py:67-100"""
        pt = np.zeros(3) if point_in_global is None else point_in_global
        self.external_forces[segment_index].append((FEXT_GLOBAL, self._param(external_force, pt)))

    def add_in_global_local_point(self, segment_index: int, external_force, point_in_local=None):
        """external_force_global_local_point.py:68-104"""
        pt = np.zeros(3) if point_in_local is None else point_in_local
        self.external_forces[segment_index].append((FEXT_GLOBAL_LOCAL_POINT, self._param(external_force, pt)))

    def add_in_global_on_proximal(self, segment_index: int, external_force):
        """external_force_global_on_proximal.py:65-115"""
        self.external_forces[segment_index].append((FEXT_GLOBAL_ON_PROXIMAL, self._param(external_force)))

    def add_in_local(self, segment_index: int, external_force, point_in_local=None, transformation_matrix=None):
        """external_force_in_local.py:85-131; ``transformation_matrix`` as in the reference (its
        inverse-transpose rotates segment-frame vectors to the global frame through ``[u v w]``)."""
        if transformation_matrix is None:
            raise ValueError("transformation_matrix is required for forces expressed in the local frame")
        pt = np.zeros(3) if point_in_local is None else point_in_local
        tinv = np.linalg.inv(np.asarray(transformation_matrix, dtype=np.float64).T)
        self.external_forces[segment_index].append((FEXT_LOCAL, self._param(external_force, pt, tinv)))

    def __iter__(self):
        return iter(self.external_forces)

    def __len__(self):
        return len(self.external_forces)

    # ------------------------------------------------------------------ lowering
    @classmethod
    def from_reference(cls, fext) -> "ExternalForceSet":
        """Accept a ``bionc_numpy.ExternalForceSet`` (duck-typed by the force class names)."""
        out = cls.empty_from_nb_segment(len(fext.external_forces))
        for s, forces in enumerate(fext.external_forces):
            for f in forces:
                if isinstance(f, tuple):
                    out.external_forces[s].append(f)
                    continue
                name = type(f).__name__
                if name == "ExternalForceInGlobalOnProximal":
                    out.add_in_global_on_proximal(s, f.external_forces)
                elif name == "ExternalForceInGlobalLocalPoint":
                    out.add_in_global_local_point(s, f.external_forces, f.application_point_in_local)
                elif name == "ExternalForceInGlobal":
                    out.add_in_global(s, f.external_forces, f.application_point_in_global)
                elif name == "ExternalForceInLocal":
                    out.external_forces[s].append(
                        (FEXT_LOCAL, cls._param(f.external_forces, f.application_point_in_local, f.transformation_matrix_inv))
                    )
                else:
                    raise TypeError(f"unknown external force class {name}")
        return out

    def to_arrays(self):
        seg, kind, par = [], [], []
        for s, forces in enumerate(self.external_forces):
            for k, p in forces:
                seg.append(s), kind.append(k), par.append(p)
        return (
            np.ascontiguousarray(seg, dtype=np.int32),
            np.ascontiguousarray(kind, dtype=np.int32),
            np.ascontiguousarray(np.array(par, dtype=np.float64).reshape(len(seg), FEXT_NPARAM)),
        )
