"""Single-segment view of a compiled model: the dynamics-side methods of ``bionc_numpy.NaturalSegment``
(``bionc_numpy/natural_segment.py``) on the CUDA backend, batched like the model-level API.

A segment on its own is a one-segment model without joints; the view compiles that sub-table once and runs
the same kernels (``differential_algebraic_equation`` = the 18x18 system of ``natural_segment.py:574-633``
is the N = 1, m = 6 case of ``forward_dynamics``)."""

from __future__ import annotations

import numpy as np

from .model_table import BLK_NPARAM, ModelTable


def segment_table(table: ModelTable, i: int) -> ModelTable:
    """The one-segment, joint-free sub-table of segment ``i`` (its markers included)."""
    keep = np.nonzero(table.marker_segment == i)[0] if table.nb_markers else np.zeros(0, dtype=int)
    return ModelTable(
        segment_names=[table.segment_names[i]],
        seg_phi_const=table.seg_phi_const[i : i + 1].copy(),
        seg_geometry=table.seg_geometry[i : i + 1].copy(),
        seg_mass=table.seg_mass[i : i + 1].copy(),
        seg_com=table.seg_com[i : i + 1].copy(),
        seg_J=table.seg_J[i : i + 1].copy(),
        seg_rigid_row=np.zeros(1, dtype=np.int32),
        joint_names=[],
        joint_kinds=[],
        blk_type=np.zeros(0, dtype=np.int32),
        blk_parent=np.zeros(0, dtype=np.int32),
        blk_child=np.zeros(0, dtype=np.int32),
        blk_joint=np.zeros(0, dtype=np.int32),
        blk_row_h=np.zeros(0, dtype=np.int32),
        blk_row_j=np.zeros(0, dtype=np.int32),
        blk_param=np.zeros((0, BLK_NPARAM)),
        marker_names=[table.marker_names[k] for k in keep],
        marker_segment=np.zeros(len(keep), dtype=np.int32),
        marker_position=table.marker_position[keep].copy() if len(keep) else np.zeros((0, 3)),
        marker_technical=table.marker_technical[keep].copy() if len(keep) else np.zeros(0, dtype=np.int32),
    ).validate()


class NaturalSegment:
    """``model.segment(i)``.  Inputs are ``Qi`` / ``Qdoti`` of shape (12,), (12, 1) or (B, 12)."""

    def __init__(self, model, index: int):
        if not 0 <= index < model.nb_segments:
            raise IndexError(f"segment index {index} outside [0, {model.nb_segments})")
        self.index = index
        self.name = model.table.segment_names[index]
        self._parent = model
        self._sub = None

    @property
    def _model(self):
        if self._sub is None:
            self._sub = type(self._parent).from_table(segment_table(self._parent.table, self.index), device=self._parent.device)
        return self._sub

    # ------------------------------------------------------------------ constants
    @property
    def length(self) -> float:
        return float(self._parent.table.seg_geometry[self.index, 0])

    @property
    def alpha(self) -> float:
        return float(self._parent.table.seg_geometry[self.index, 1])

    @property
    def beta(self) -> float:
        return float(self._parent.table.seg_geometry[self.index, 2])

    @property
    def gamma(self) -> float:
        return float(self._parent.table.seg_geometry[self.index, 3])

    @property
    def mass(self) -> float:
        return float(self._parent.table.seg_mass[self.index])

    @property
    def natural_center_of_mass(self) -> np.ndarray:
        return self._parent.table.seg_com[self.index].copy()

    @property
    def mass_matrix(self) -> np.ndarray:
        """``G_i = M4 (x) I3`` (12, 12), natural_inertial_parameters.py:136-179"""
        i = self.index
        return self._parent.mass_matrix[12 * i : 12 * i + 12, 12 * i : 12 * i + 12].copy()

    def gravity_force(self) -> np.ndarray:
        """natural_segment.py:562-572 -> (12,)"""
        i = self.index
        return np.asarray(self._parent.gravity_forces()).reshape(-1)[12 * i : 12 * i + 12].copy()

    @property
    def nb_markers(self) -> int:
        return self._model.nb_markers

    @property
    def marker_names(self) -> list[str]:
        return self._model.marker_names

    # ------------------------------------------------------------------ state-dependent quantities
    def rigid_body_constraint(self, Qi):
        """natural_segment.py:429-448 -> (6,) / (B, 6)"""
        return self._model.rigid_body_constraints(Qi)

    def rigid_body_constraint_jacobian(self, Qi):
        """natural_segment.py:450-483 -> (6, 12) / (B, 6, 12)"""
        return self._model.rigid_body_constraints_jacobian(Qi)

    def rigid_body_constraint_acceleration_bias(self, Qdoti):
        """natural_segment.py:508-545 -> (6, 1) / (B, 6)"""
        return self._model.holonomic_constraints_acceleration_bias(Qdoti)

    def markers(self, Qi):
        """natural_segment.py:764-765 -> (3, nb_markers, 1) / (B, 3, nb_markers)"""
        return self._model.markers(Qi)

    def center_of_mass_position(self, Qi):
        """natural_segment.py:551-560 -> (3,) / (B, 3)"""
        out = self._model.center_of_mass_position(Qi)  # (3, 1, 1) for one system, (B, 3, 1) for a batch
        return out[:, :, 0] if self._is_batch(Qi) else out[:, 0, 0]

    @staticmethod
    def _is_batch(x) -> bool:
        shape = tuple(x.shape) if hasattr(x, "shape") else np.asarray(x).shape
        return len(shape) == 2 and not (shape[1] == 1 and shape[0] == 12)

    def kinetic_energy(self, Qdoti):
        """natural_segment.py:791-806"""
        return self._model.kinetic_energy(Qdoti)

    def potential_energy(self, Qi):
        """natural_segment.py:775-789 (as coded: no factor g)"""
        return self._model.potential_energy(Qi)

    def differential_algebraic_equation(self, Qi, Qdoti, stabilization: dict = None, return_status: bool = False):
        """natural_segment.py:574-633: ``[G_i Kr^T; Kr 0] [Qddot_i; lambda_i] = [gravity_i; -bias_i]`` ->
        ``(Qddot_i, lambda_i)`` of shapes (12,), (6,) for one system like the reference, (B, 12), (B, 6) for a batch.
        ``stabilization`` applies the Baumgarte terms the reference intends (its own branch at :616-618 raises a
        shape error for every input, so there is no reference output to compare with)."""
        res = self._model.forward_dynamics(Qi, Qdoti, stabilization=stabilization, return_status=return_status)
        if not self._is_batch(Qi):
            res = tuple(r[..., 0] if (r.ndim == 2 and r.shape[-1] == 1) else r for r in res)
        return res
