"""
Compiled model table: the flat, numeric description of a natural-coordinate model that the CUDA
kernels (and the CPU oracle) consume.

A bioNC model is a graph of live Python objects (``NaturalSegment``, ``Joint.*``, ``GroundJoint.*``)
pickled with dill (reference ``bionc/protocols/biomechanical_model.py:251-281``).  The GPU backend
never parses that pickle; it walks an already-loaded ``bionc_numpy`` model *once* (``from_numpy``,
the analogue of the reference's ``to_mx()`` at ``bionc/bionc_numpy/biomechanical_model.py:43-75``)
and lowers it to a handful of arrays:

* per segment: the rigid-body constants, the natural inertial parameters ``(m, c, J)`` and the 4x4
  matrix ``M4`` with ``G_i = M4 (x) I3``
  (``bionc/bionc_numpy/natural_inertial_parameters.py:153-177``);
* per *constraint block*: every joint is lowered to one or more blocks of four primitive kinds
  (``LIN3``, ``DOT``, ``CLEN``, ``SOP``), each with <= 13 doubles of parameters.  All points and
  directions are 4-coefficient linear forms over a segment's ``(u, rp, rd, w)``
  (``bionc/bionc_numpy/natural_vector.py:54-61, :128-137``);
* two row orders: the *holonomic* order of ``holonomic_constraints*`` / ``forward_dynamics``
  (``bionc/bionc_numpy/biomechanical_model.py:154-258``: per segment, joints whose child is that
  segment, then its 6 rigid rows) and the *joint-index* order of ``joint_constraints*``
  (``bionc/bionc_numpy/biomechanical_model_joints.py:13-77``).

The table is also the on-disk format of the backend (``save``/``load``: a plain ``.npz``), so a
compiled model travels to machines where the reference package is not installed.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# primitive constraint-block kinds (shared with csrc/model.cuh and oracle/bionc_oracle.py)
BLK_LIN3 = 0  # 3 rows: c0 + a_p (x) Q_p - a_c (x) Q_c            (linear, constant Jacobian, zero bias)
BLK_DOT = 1  # 1 row : D_p . D_c - cos(theta)                     (D_p a segment direction or a ground axis)
BLK_CLEN = 2  # 1 row : |P_p - P_c|^2 - l^2
BLK_SOP = 3  # 1 row : (P - A) . n - r, Jacobian blocks as coded in the reference (parent/child swapped)

BLK_ROWS = {BLK_LIN3: 3, BLK_DOT: 1, BLK_CLEN: 1, BLK_SOP: 1}
BLK_NPARAM = 16

# external-force kinds (reference bionc/bionc_numpy/external_force*.py)
FEXT_GLOBAL_ON_PROXIMAL = 0
FEXT_GLOBAL_LOCAL_POINT = 1
FEXT_GLOBAL = 2
FEXT_LOCAL = 3
FEXT_JOINT_REACTION = 4  # reaction of a joint generalized force on the joint's parent (generalized_force.py:98-129)
FEXT_NPARAM = 24  # [torque(3), force(3), point(3), Tinv(9)] padded


class UnsupportedModelError(NotImplementedError):
    """Raised at model-compile time for models the reference itself cannot simulate forward."""


def _coefs_from_interpolation_matrix(N) -> np.ndarray:
    """A 3x12 interpolation matrix of the reference is always ``coefs (x) I3``; return the 4 coefs."""
    N = np.asarray(N, dtype=np.float64)
    if N.shape != (3, 12):
        raise ValueError(f"interpolation matrix must be 3x12, got {N.shape}")
    coefs = N[0, [0, 3, 6, 9]].copy()
    if not np.array_equal(N, np.kron(coefs[None, :], np.eye(3))):
        raise ValueError("interpolation matrix is not of the form coefs (x) I3")
    return coefs


def _direction_coefs(natural_vector) -> np.ndarray:
    """``NaturalVector.interpolate().rot`` (natural_vector.py:54-61, :128-137) as 4 coefficients."""
    a = np.asarray(natural_vector, dtype=np.float64).reshape(3)
    return np.array([a[0], a[1], -a[1], a[2]])


def _point_coefs(natural_vector) -> np.ndarray:
    """``NaturalVector.interpolate()`` (natural_vector.py:54-61) as 4 coefficients."""
    a = np.asarray(natural_vector, dtype=np.float64).reshape(3)
    return np.array([a[0], 1.0 + a[1], -a[1], a[2]])


def m4_from_natural_parameters(mass, com, J) -> np.ndarray:
    """``M4`` such that ``G_i = M4 (x) I3``; same entries as
    ``NaturalInertialParameters._update_mass_matrix`` (natural_inertial_parameters.py:153-177).
    Broadcasts over leading axes."""
    mass = np.asarray(mass, dtype=np.float64)
    com = np.asarray(com, dtype=np.float64)
    J = np.asarray(J, dtype=np.float64)
    M4 = np.zeros(mass.shape + (4, 4))
    M4[..., 0, 0] = J[..., 0, 0]
    M4[..., 0, 1] = mass * com[..., 0] + J[..., 0, 1]
    M4[..., 0, 2] = -J[..., 0, 1]
    M4[..., 0, 3] = J[..., 0, 2]
    M4[..., 1, 1] = mass + 2 * mass * com[..., 1] + J[..., 1, 1]
    M4[..., 1, 2] = -(mass * com[..., 1] + J[..., 1, 1])
    M4[..., 1, 3] = mass * com[..., 2] + J[..., 1, 2]
    M4[..., 2, 2] = J[..., 1, 1]
    M4[..., 2, 3] = -J[..., 1, 2]
    M4[..., 3, 3] = J[..., 2, 2]
    iu = np.triu_indices(4, 1)
    M4[..., iu[1], iu[0]] = M4[..., iu[0], iu[1]]
    return M4


def natural_inertial_parameters_from_cartesian(mass, center_of_mass, inertia, length, alpha, beta, gamma):
    """Cartesian inertial parameters given in the segment coordinate system (``Buv`` convention) ->
    natural ``(c_nat, J)``, exactly as the reference codes it
    (``NaturalSegment.with_cartesian_inertial_parameters``, natural_segment.py:223-264, and
    ``NaturalInertialParameters.compute_natural_center_of_mass`` / ``compute_pseudo_inertia_matrix``,
    natural_inertial_parameters.py:298-352):

    * the matrix handed over is the TRANSPOSE of ``transformation_matrix.py:7-33``'s ``Buv``;
    * the middle block is ``I + m (c.c) E - (c.c) 11^T`` -- ``np.dot(c.T, c)`` is the scalar ``c.c``, the
      ``m c c^T`` outer product of the textbook formula is never formed.
    """
    c = np.asarray(center_of_mass, dtype=np.float64).reshape(3)
    I = np.asarray(inertia, dtype=np.float64).reshape(3, 3)
    ca, cb, cg, sg = np.cos(alpha), np.cos(beta), np.cos(gamma), np.sin(gamma)
    Buv = np.array(
        [
            [1, length * cg, cb],
            [0, length * sg, (ca - cb * cg) / sg],
            [0, 0, np.sqrt(1 - cb**2 - ((ca - cb * cg) / sg) ** 2)],
        ]
    )
    Binv = np.linalg.inv(Buv.T)
    cc = float(c @ c)
    middle = I + mass * cc * np.eye(3) - cc
    return Binv @ c, Binv @ (middle @ Binv.T)


def gravity_from_natural_parameters(mass, com) -> np.ndarray:
    """``m N_com^T (0, 0, -9.81)`` (natural_segment.py:562-572).  Broadcasts over leading axes."""
    mass = np.asarray(mass, dtype=np.float64)
    com = np.asarray(com, dtype=np.float64)
    f = np.zeros(mass.shape + (12,))
    g = -9.81
    f[..., 2] = com[..., 0] * mass * g
    f[..., 5] = (1 + com[..., 1]) * mass * g
    f[..., 8] = -com[..., 1] * mass * g
    f[..., 11] = com[..., 2] * mass * g
    return f


@dataclass
class ModelTable:
    """Flat numeric model.  Every array is float64 or int32 and C-contiguous."""

    segment_names: list[str]
    # rigid-body constants, evaluated exactly as the reference does (natural_segment.py:441-446):
    # [L*cos(gamma), cos(beta), L**2, L*cos(alpha)]
    seg_phi_const: np.ndarray  # (N, 4)
    seg_geometry: np.ndarray  # (N, 4) [length, alpha, beta, gamma] (informational)
    seg_mass: np.ndarray  # (N,)
    seg_com: np.ndarray  # (N, 3) natural centre of mass
    seg_J: np.ndarray  # (N, 3, 3) natural pseudo-inertia
    seg_rigid_row: np.ndarray  # (N,) int32: first holonomic row of the segment's 6 rigid rows
    # constraint blocks
    joint_names: list[str] = field(default_factory=list)
    joint_kinds: list[str] = field(default_factory=list)  # reference class name per joint (with constraints)
    blk_type: np.ndarray = None  # (nblk,) int32
    blk_parent: np.ndarray = None  # (nblk,) int32, -1 = ground
    blk_child: np.ndarray = None  # (nblk,) int32
    blk_joint: np.ndarray = None  # (nblk,) int32 index into joint_names
    blk_row_h: np.ndarray = None  # (nblk,) int32 first row in holonomic order
    blk_row_j: np.ndarray = None  # (nblk,) int32 first row in joint-index order
    blk_param: np.ndarray = None  # (nblk, BLK_NPARAM) float64
    # markers (biomechanical_model_markers.py:32-57): model order = segment by segment; optional
    marker_names: list[str] = field(default_factory=list)
    marker_segment: np.ndarray = None  # (nmk,) int32
    marker_position: np.ndarray = None  # (nmk, 3) natural coordinates (n_u, n_v, n_w) of the marker in its segment
    marker_technical: np.ndarray = None  # (nmk,) int32 0/1
    # every joint of the model in the reference's joint order (model.joints, free joints included): what joint
    # generalized forces are indexed by (generalized_force.py:232-311); optional
    all_joint_names: list[str] = field(default_factory=list)
    all_joint_parent: np.ndarray = None  # (nj,) int32, -1 = ground
    all_joint_child: np.ndarray = None  # (nj,) int32

    # ------------------------------------------------------------------ counts (protocols/biomechanical_model.py:347-453)
    @property
    def nb_segments(self) -> int:
        return len(self.segment_names)

    @property
    def nb_Q(self) -> int:
        return 12 * self.nb_segments

    @property
    def nb_blocks(self) -> int:
        return int(self.blk_type.shape[0])

    @property
    def nb_markers(self) -> int:
        return 0 if self.marker_segment is None else int(self.marker_segment.shape[0])

    @property
    def nb_rigid_body_constraints(self) -> int:
        return 6 * self.nb_segments

    @property
    def nb_joint_constraints(self) -> int:
        return int(sum(BLK_ROWS[int(t)] for t in self.blk_type))

    @property
    def nb_holonomic_constraints(self) -> int:
        return self.nb_rigid_body_constraints + self.nb_joint_constraints

    # ------------------------------------------------------------------ derived constants
    @property
    def seg_M4(self) -> np.ndarray:
        return m4_from_natural_parameters(self.seg_mass, self.seg_com, self.seg_J)

    @property
    def seg_gravity(self) -> np.ndarray:
        return gravity_from_natural_parameters(self.seg_mass, self.seg_com)

    @property
    def mass_matrix(self) -> np.ndarray:
        """Block-diagonal ``G`` (biomechanical_model.py:77-96)."""
        N = self.nb_segments
        G = np.zeros((12 * N, 12 * N))
        M4 = self.seg_M4
        for i in range(N):
            G[12 * i : 12 * i + 12, 12 * i : 12 * i + 12] = np.kron(M4[i], np.eye(3))
        return G

    @property
    def is_mass_matrix_positive_definite(self) -> bool:
        return bool(np.all(np.linalg.eigvalsh(self.seg_M4) > 0))

    # ------------------------------------------------------------------ io
    _ARRAYS = (
        "seg_phi_const seg_geometry seg_mass seg_com seg_J seg_rigid_row blk_type blk_parent blk_child "
        "blk_joint blk_row_h blk_row_j blk_param"
    ).split()

    def validate(self) -> "ModelTable":
        N = self.nb_segments
        for name in self._ARRAYS:
            a = getattr(self, name)
            want = np.int32 if name.startswith("blk_") and name != "blk_param" or name == "seg_rigid_row" else np.float64
            setattr(self, name, np.ascontiguousarray(a, dtype=want))
        nmk = len(self.marker_names)
        self.marker_segment = np.ascontiguousarray(self.marker_segment if self.marker_segment is not None else np.zeros(0), dtype=np.int32)
        self.marker_position = np.ascontiguousarray(
            self.marker_position if self.marker_position is not None else np.zeros((0, 3)), dtype=np.float64).reshape(-1, 3)
        self.marker_technical = np.ascontiguousarray(self.marker_technical if self.marker_technical is not None else np.zeros(0), dtype=np.int32)
        assert self.marker_segment.shape == (nmk,) and self.marker_position.shape == (nmk, 3) and self.marker_technical.shape == (nmk,)
        if nmk:
            assert self.marker_segment.min() >= 0 and self.marker_segment.max() < N
        nj = len(self.all_joint_names)
        self.all_joint_parent = np.ascontiguousarray(self.all_joint_parent if self.all_joint_parent is not None else np.zeros(0), dtype=np.int32)
        self.all_joint_child = np.ascontiguousarray(self.all_joint_child if self.all_joint_child is not None else np.zeros(0), dtype=np.int32)
        assert self.all_joint_parent.shape == (nj,) and self.all_joint_child.shape == (nj,)
        assert self.seg_phi_const.shape == (N, 4) and self.seg_J.shape == (N, 3, 3)
        assert self.blk_param.shape == (self.nb_blocks, BLK_NPARAM)
        if self.nb_blocks:
            assert self.blk_child.min() >= 0 and self.blk_child.max() < N
            assert self.blk_parent.min() >= -1 and self.blk_parent.max() < N
        # every holonomic row is claimed exactly once
        rows = np.zeros(self.nb_holonomic_constraints, dtype=int)
        for i in range(N):
            rows[self.seg_rigid_row[i] : self.seg_rigid_row[i] + 6] += 1
        for b in range(self.nb_blocks):
            rows[self.blk_row_h[b] : self.blk_row_h[b] + BLK_ROWS[int(self.blk_type[b])]] += 1
        assert np.all(rows == 1), "holonomic row map is not a permutation"
        return self

    def save(self, filename: str) -> None:
        np.savez(
            filename,
            format_version=np.int32(3),
            all_joint_names=np.array(self.all_joint_names, dtype=str),
            all_joint_parent=self.all_joint_parent,
            all_joint_child=self.all_joint_child,
            segment_names=np.array(self.segment_names),
            marker_names=np.array(self.marker_names, dtype=str),
            marker_segment=self.marker_segment,
            marker_position=self.marker_position,
            marker_technical=self.marker_technical,
            joint_names=np.array(self.joint_names),
            joint_kinds=np.array(self.joint_kinds),
            **{k: getattr(self, k) for k in self._ARRAYS},
        )

    @classmethod
    def load(cls, filename: str) -> "ModelTable":
        with np.load(filename, allow_pickle=False) as z:
            version = int(z["format_version"])
            if version not in (1, 2, 3):  # 2 = 1 + the marker arrays, 3 = 2 + the joint list
                raise ValueError(f"unknown model table version {version}")
            kw = {k: z[k] for k in cls._ARRAYS}
            if version >= 2:
                kw.update(marker_names=[str(s) for s in z["marker_names"]], marker_segment=z["marker_segment"],
                          marker_position=z["marker_position"], marker_technical=z["marker_technical"])
            if version >= 3:
                kw.update(all_joint_names=[str(s) for s in z["all_joint_names"]], all_joint_parent=z["all_joint_parent"],
                          all_joint_child=z["all_joint_child"])
            return cls(
                segment_names=[str(s) for s in z["segment_names"]],
                joint_names=[str(s) for s in z["joint_names"]],
                joint_kinds=[str(s) for s in z["joint_kinds"]],
                **kw,
            ).validate()

    # ------------------------------------------------------------------ the model compiler
    @classmethod
    def from_numpy(cls, model) -> "ModelTable":
        """Lower a loaded ``bionc.bionc_numpy.BiomechanicalModel`` to a table.

        Row orders are taken from the reference's own bookkeeping
        (``joints_from_child_index(i, remove_free_joints=True)``, ``constraints_index``) so they are
        the reference's by construction.  Raises the reference's own exception types for models
        its ``forward_dynamics`` cannot run (SURVEY.md section 8(b) error conventions).
        """
        segments = list(model.segments_no_ground.values())
        N = len(segments)
        for pos, s in enumerate(segments):
            if s.index != pos:
                raise ValueError(f"segment {s.name!r} has index {s.index}, expected {pos}")

        names, phi_const, geom, mass, com, J = [], np.zeros((N, 4)), np.zeros((N, 4)), [], [], []
        for i, s in enumerate(segments):
            names.append(str(s.name))
            L, al, be, ga = float(s.length), float(s.alpha), float(s.beta), float(s.gamma)
            geom[i] = (L, al, be, ga)
            # same expressions as natural_segment.py:441-446 (np.cos(pi/2) = 6.1e-17, kept)
            phi_const[i] = (L * np.cos(ga), np.cos(be), L**2, L * np.cos(al))
            Gi = s.mass_matrix
            if s.mass is None or s.natural_center_of_mass is None or Gi is None:
                # reference: mass_matrix is None (biomechanical_model.py:89-92) and augmented_mass_matrix raises
                raise ValueError(
                    f"segment {s.name!r} has no inertial parameters: the reference mass matrix is None "
                    "and forward dynamics cannot be computed"
                )
            mass.append(float(s.mass))
            com.append(np.asarray(s.natural_center_of_mass, dtype=np.float64).reshape(3))
            # The pseudo-inertia is read back from the segment's own G_i = M4 (x) I3 -- the matrix the
            # reference dynamics actually uses -- because NaturalSegment.set_inertial_parameters leaves
            # `natural_pseudo_inertia` unset (natural_segment.py:177-193).  Inverse of
            # natural_inertial_parameters.py:153-177: only J00, J01, J02, J11, J12, J22 enter G_i.
            M4 = np.asarray(Gi, dtype=np.float64)[::3, ::3]
            Ji = np.zeros((3, 3))
            Ji[0, 0], Ji[0, 1], Ji[0, 2] = M4[0, 0], -M4[0, 2], M4[0, 3]
            Ji[1, 1], Ji[1, 2], Ji[2, 2] = M4[2, 2], -M4[2, 3], M4[3, 3]
            J.append(np.triu(Ji) + np.triu(Ji, 1).T)

        blocks = []  # dicts
        joint_names, joint_kinds = [], []
        seg_rigid_row = np.zeros(N, dtype=np.int32)
        row_h = 0
        for i in range(N):
            for j in model.joints_from_child_index(i, remove_free_joints=True):
                kind = type(j).__qualname__
                jrow = int(model.joints.constraints_index(j.index).start)
                jid = len(joint_names)
                joint_names.append(str(j.name))
                joint_kinds.append(kind)
                new = _lower_joint(j, kind)
                r = 0
                for blk in new:
                    blk.update(
                        parent=-1 if j.parent is None else int(j.parent.index),
                        child=int(j.child.index),
                        joint=jid,
                        row_h=row_h + r,
                        row_j=jrow + r,
                    )
                    r += BLK_ROWS[blk["type"]]
                    blocks.append(blk)
                if r != j.nb_constraints:
                    raise AssertionError(f"joint {j.name}: lowered {r} rows, reference has {j.nb_constraints}")
                row_h += r
            seg_rigid_row[i] = row_h
            row_h += 6

        # markers in model order (biomechanical_model_markers.py: indexes() walks the segments in order)
        mk_names, mk_seg, mk_pos, mk_tech = [], [], [], []
        for i, s in enumerate(segments):
            for mk in s._markers:
                mk_names.append(str(mk.name)), mk_seg.append(i)
                mk_pos.append(np.asarray(mk.position, dtype=np.float64).reshape(3))
                mk_tech.append(1 if mk.is_technical else 0)

        nb = len(blocks)
        param = np.zeros((nb, BLK_NPARAM))
        for b, blk in enumerate(blocks):
            p = np.asarray(blk["param"], dtype=np.float64)
            param[b, : p.size] = p
        table = cls(
            segment_names=names,
            seg_phi_const=phi_const,
            seg_geometry=geom,
            seg_mass=np.array(mass),
            seg_com=np.array(com).reshape(N, 3),
            seg_J=np.array(J).reshape(N, 3, 3),
            seg_rigid_row=seg_rigid_row,
            joint_names=joint_names,
            joint_kinds=joint_kinds,
            blk_type=np.array([b["type"] for b in blocks], dtype=np.int32),
            blk_parent=np.array([b["parent"] for b in blocks], dtype=np.int32),
            blk_child=np.array([b["child"] for b in blocks], dtype=np.int32),
            blk_joint=np.array([b["joint"] for b in blocks], dtype=np.int32),
            blk_row_h=np.array([b["row_h"] for b in blocks], dtype=np.int32),
            blk_row_j=np.array([b["row_j"] for b in blocks], dtype=np.int32),
            blk_param=param,
            marker_names=mk_names,
            marker_segment=np.array(mk_seg, dtype=np.int32),
            marker_position=np.array(mk_pos, dtype=np.float64).reshape(len(mk_seg), 3),
            marker_technical=np.array(mk_tech, dtype=np.int32),
            all_joint_names=[str(j.name) for j in model.joints.values()],
            all_joint_parent=np.array([-1 if j.parent is None else int(j.parent.index) for j in model.joints.values()], dtype=np.int32),
            all_joint_child=np.array([int(j.child.index) for j in model.joints.values()], dtype=np.int32),
        ).validate()
        if row_h != model.nb_holonomic_constraints:
            raise AssertionError("holonomic row count differs from the reference model")
        # cross-check the derived mass matrix with the reference's cached one (biomechanical_model.py:77-96)
        G_ref = model.mass_matrix
        if G_ref is not None and not np.allclose(G_ref, table.mass_matrix, rtol=0, atol=1e-13 * max(1.0, np.abs(G_ref).max())):
            raise AssertionError("mass matrix rebuilt from (m, c, J) differs from the reference model's")
        return table


_E_RP = np.array([0.0, 1.0, 0.0, 0.0])
_E_RD = np.array([0.0, 0.0, 1.0, 0.0])
_ZERO4 = np.zeros(4)


def _lin3(a_p, a_c, c0=(0.0, 0.0, 0.0)):
    return dict(type=BLK_LIN3, param=np.concatenate([a_p, a_c, np.asarray(c0, dtype=np.float64).reshape(3)]))


def _dot(a_p, a_c, cos_theta):
    return dict(type=BLK_DOT, param=np.concatenate([a_p, a_c, [cos_theta]]))


def _cart_axis(vec) -> np.ndarray:
    """ground axis as a padded 4-vector [x, y, z, 0] (cartesian_vector.py:30-39)."""
    v = np.asarray(vec, dtype=np.float64).reshape(3)
    return np.array([v[0], v[1], v[2], 0.0])


def _lower_joint(j, kind: str) -> list[dict]:
    """One reference joint -> primitive blocks, in the joint's own row order."""
    if kind == "GroundJoint.Hinge":  # joints_with_ground.py:142-210
        out = [_lin3(_ZERO4, _E_RP)]
        for k in range(2):
            out.append(_dot(_cart_axis(j.parent_vector[k]), _direction_coefs(j.child_vector[k]), np.cos(j.theta[k])))
        return out
    if kind == "GroundJoint.Universal":  # joints_with_ground.py:293-354
        return [
            _lin3(_ZERO4, _E_RP),
            _dot(_cart_axis(j.parent_vector), _direction_coefs(j.child_vector), np.cos(j.theta)),
        ]
    if kind == "GroundJoint.Spherical":  # joints_with_ground.py:401-460
        p = np.zeros(3) if j.ground_application_point is None else np.asarray(j.ground_application_point, float).reshape(3)
        return [_lin3(_ZERO4, _E_RP, p)]
    if kind == "GroundJoint.Weld":  # joints_with_ground.py:505-528
        return [
            _lin3(_ZERO4, _E_RP, np.asarray(j.rp_child_ref, float).reshape(3)),
            _lin3(_ZERO4, _E_RD, np.asarray(j.rd_child_ref, float).reshape(3)),
        ]
    if kind == "Joint.Hinge":  # joints.py:190-282
        out = [_lin3(_E_RD, _E_RP)]
        for k in range(2):
            out.append(
                _dot(_direction_coefs(j.parent_vector[k]), _direction_coefs(j.child_vector[k]), np.cos(j.theta[k]))
            )
        return out
    if kind == "Joint.Universal":  # joints.py:369-453
        return [
            _lin3(_E_RD, _E_RP),
            _dot(_direction_coefs(j.parent_vector), _direction_coefs(j.child_vector), np.cos(j.theta)),
        ]
    if kind == "Joint.Spherical":  # joints.py:523-587
        return [
            _lin3(
                _coefs_from_interpolation_matrix(j.parent_point.interpolation_matrix),
                _coefs_from_interpolation_matrix(j.child_point.interpolation_matrix),
            )
        ]
    if kind == "Joint.ConstantLength":  # joints.py:1300-1382
        return [
            dict(
                type=BLK_CLEN,
                param=np.concatenate(
                    [
                        _coefs_from_interpolation_matrix(j.parent_point.interpolation_matrix),
                        _coefs_from_interpolation_matrix(j.child_point.interpolation_matrix),
                        [float(j.length) ** 2],
                    ]
                ),
            )
        ]
    if kind == "Joint.SphereOnPlane":  # joints.py:657-748
        return [
            dict(
                type=BLK_SOP,
                param=np.concatenate(
                    [
                        _coefs_from_interpolation_matrix(j.sphere_center.interpolation_matrix),
                        _coefs_from_interpolation_matrix(j.plane_point.interpolation_matrix),
                        _coefs_from_interpolation_matrix(j.plane_normal.interpolation_matrix),
                        [float(j.sphere_radius)],
                    ]
                ),
            )
        ]
    if kind == "Joint.Free":
        # reference: forward_dynamics raises AttributeError (no constraint_acceleration_bias), SURVEY a14
        raise AttributeError("'Free' object has no attribute 'constraint_acceleration_bias'")
    if kind in ("Joint.EllipsoidOnPlane", "Joint.PointOnEllipsoid", "Joint.TwoPointsOnEllipsoid"):
        # joints.py:946, :1090, :1229
        raise UnsupportedModelError(f"{kind}: constraint_acceleration_bias is not implemented in the reference")
    raise UnsupportedModelError(f"unknown joint class {kind}")
