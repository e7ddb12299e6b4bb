"""
Synthetic, constraint-consistent initial states for batches of independent systems (the workloads
of BASELINE.json / SURVEY.md section 8(d)).  Pure numpy, works from a ``ModelTable`` only.

Natural coordinates of one segment: ``Q_i = (u, rp, rd, w)`` with ``v = rp - rd``
(reference ``bionc/bionc_numpy/natural_coordinates.py:51-69``).  A rigid placement of a segment
with geometry ``(L, alpha, beta, gamma)`` is ``[u v w] = R @ B`` where the columns of ``B`` are
``u, v, w`` in the segment's orthonormal frame (the ``Buv`` convention of
``bionc/bionc_numpy/transformation_matrix.py:7-33``, obtained here as the Cholesky factor of the
Gram matrix), and a rigid velocity is ``(w x u, vp, vp - w x v, w x w_axis)``.
"""

from __future__ import annotations

import numpy as np

from ..bionc_cuda.model_table import BLK_LIN3, ModelTable


def nominal_frame(seg_phi_const_row) -> np.ndarray:
    """3x3 matrix whose columns are ``u, v, w`` of an undeformed segment: ``u`` along x, ``v`` in
    the xy-plane (``Buv``).  Built from the same constants the rigid-body constraints use, so the
    constraints are satisfied to rounding."""
    Lcg, cb, L2, Lca = seg_phi_const_row
    gram = np.array([[1.0, Lcg, cb], [Lcg, L2, Lca], [cb, Lca, 1.0]])
    return np.linalg.cholesky(gram).T


def rotation_from_rotvec(rv: np.ndarray) -> np.ndarray:
    """Rodrigues formula, vectorised over leading axes: rv (..., 3) -> R (..., 3, 3)."""
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv, axis=-1)
    safe = np.where(th > 0, th, 1.0)
    k = rv / safe[..., None]
    K = np.zeros(rv.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None, None], np.cos(th)[..., None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def _cross(a, b):
    return np.cross(a, b)


def _place(R, frame, rp, omega, vp):
    """Segment coordinates/velocities from orientation R (B,3,3), origin rp (B,3), angular
    velocity omega (B,3) and origin velocity vp (B,3)."""
    uvw = R @ frame  # (B,3,3) columns u v w
    u, v, w = uvw[..., 0], uvw[..., 1], uvw[..., 2]
    Q = np.concatenate([u, rp, rp - v, w], axis=-1)
    Qd = np.concatenate([_cross(omega, u), vp, vp - _cross(omega, v), _cross(omega, w)], axis=-1)
    return Q, Qd


def _point(coefs, Qi):
    return coefs[0] * Qi[..., 0:3] + coefs[1] * Qi[..., 3:6] + coefs[2] * Qi[..., 6:9] + coefs[3] * Qi[..., 9:12]


def tree_states(
    table: ModelTable,
    batch: int,
    seed: int = 0,
    root_position=(0.0, 0.0, 1.0),
    root_position_sigma: float = 0.05,
    rotation_sigma: float = 0.2,
    omega_sigma: float = 1.0,
    root_velocity_sigma: float = 0.0,
):
    """Consistent ``(Q, Qdot)`` of shape ``(batch, 12 N)`` for a tree of segments connected by
    point-to-point (``LIN3``) joints -- spherical joints, ground spherical joints, free roots --
    (config 3 "lower limb" and config 5 "full body").  Each segment gets the orientation of its
    parent composed with a random rotation (sigma ``rotation_sigma`` rad) and the angular velocity
    of its parent plus a random one (sigma ``omega_sigma`` rad/s); positions and linear velocities
    are propagated through the joint points so that ``Phi = 0`` and ``K Qdot = 0`` to rounding."""
    rng = np.random.default_rng(seed)
    N = table.nb_segments
    attach = {}
    for b in range(table.nb_blocks):
        c = int(table.blk_child[b])
        if int(table.blk_type[b]) != BLK_LIN3 or c in attach:
            raise ValueError("tree_states handles one point-to-point joint per child segment only")
        attach[c] = (int(table.blk_parent[b]), table.blk_param[b, 0:4], table.blk_param[b, 4:8], table.blk_param[b, 8:11])
    Q = np.zeros((batch, 12 * N))
    Qd = np.zeros((batch, 12 * N))
    Rs, oms = {}, {}
    for i in range(N):
        frame = nominal_frame(table.seg_phi_const[i])
        dR = rotation_from_rotvec(rng.normal(0.0, rotation_sigma, (batch, 3)))
        dom = rng.normal(0.0, omega_sigma, (batch, 3))
        if i not in attach:  # free root
            R, om = dR, dom
            rp = np.asarray(root_position) + rng.normal(0.0, root_position_sigma, (batch, 3))
            vp = rng.normal(0.0, root_velocity_sigma, (batch, 3)) if root_velocity_sigma > 0 else np.zeros((batch, 3))
            Qi, Qdi = _place(R, frame, rp, om, vp)
        else:
            p, a_p, a_c, c0 = attach[i]
            if p >= i:
                raise ValueError("parent segments must come before their children")
            if p >= 0:
                R, om = Rs[p] @ dR, oms[p] + dom
                Pp = _point(a_p, Q[:, 12 * p : 12 * p + 12]) + c0
                Vp = _point(a_p, Qd[:, 12 * p : 12 * p + 12])
            else:
                R, om = dR, dom
                Pp = np.broadcast_to(c0, (batch, 3))
                Vp = np.zeros((batch, 3))
            # child point P_c = a_c0 u + (a_c1 + a_c2) rp - a_c2 v + a_c3 w must equal Pp
            s = a_c[1] + a_c[2]
            Q0, Qd0 = _place(R, frame, np.zeros((batch, 3)), om, np.zeros((batch, 3)))
            rp = (Pp - _point(a_c, Q0)) / s
            vp = (Vp - _point(a_c, Qd0)) / s
            Qi, Qdi = _place(R, frame, rp, om, vp)
        Rs[i], oms[i] = R, om
        Q[:, 12 * i : 12 * i + 12], Qd[:, 12 * i : 12 * i + 12] = Qi, Qdi
    return Q, Qd


def hinge_pendulum_states(batch: int, seed: int = 0, omega_sigma: float = 1.0, axis=(1.0, 0.0, 0.0)):
    """Config 2: ``pendulum_with_force.nmod`` -- one segment, ground hinge about the global X axis
    (``examples/forward_dynamics/pendulum_with_force.py:16-48``).  Hinge angle ~ U(-pi, pi) applied
    to the nominal pose ``u=(1,0,0), rp=0, rd=(0,-1,0), w=(0,0,1)``, rate ~ N(0, omega_sigma)."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-np.pi, np.pi, batch)
    rate = rng.normal(0.0, omega_sigma, batch)
    ax = np.asarray(axis, dtype=np.float64)
    R = rotation_from_rotvec(theta[:, None] * ax)
    om = rate[:, None] * ax
    frame = np.eye(3)  # u = x, v = y (L = 1), w = z
    return _place(R, frame, np.zeros((batch, 3)), om, np.zeros((batch, 3)))


def planar_chain_states(table: ModelTable, batch: int, seed: int = 0, angle_sigma: float = 0.5, omega_sigma: float = 1.0):
    """n-link hinge pendulum (``examples/forward_dynamics/n_link_pendulum.py:140-187``): every
    link hangs from the distal point of the previous one and rotates about the global X axis."""
    rng = np.random.default_rng(seed)
    N = table.nb_segments
    Q = np.zeros((batch, 12 * N))
    Qd = np.zeros((batch, 12 * N))
    ax = np.array([1.0, 0.0, 0.0])
    ang = np.zeros(batch)
    rate = np.zeros(batch)
    rp = np.zeros((batch, 3))
    vp = np.zeros((batch, 3))
    for i in range(N):
        ang = ang + rng.normal(0.0, angle_sigma, batch)
        rate = rate + rng.normal(0.0, omega_sigma, batch)
        frame = nominal_frame(table.seg_phi_const[i])
        Qi, Qdi = _place(rotation_from_rotvec(ang[:, None] * ax), frame, rp, rate[:, None] * ax, vp)
        Q[:, 12 * i : 12 * i + 12], Qd[:, 12 * i : 12 * i + 12] = Qi, Qdi
        rp, vp = Qi[:, 6:9], Qdi[:, 6:9]  # next link starts at this link's distal point
    return Q, Qd


def rigidly_moved_states(Q_ref: np.ndarray, batch: int, seed: int = 0, angle_max: float = 0.3, omega_sigma: float = 0.0, pivot=(0.0, 0.0, 0.0)):
    """Config 4 (knee parallel mechanism): rotate a whole consistent configuration ``Q_ref`` (12 N,)
    rigidly about ``pivot`` by an angle ~ U(-angle_max, angle_max) about a random axis; optionally
    give the whole mechanism a rigid angular velocity about the pivot.  Every closed-loop
    constraint stays satisfied because all of them are rotation invariant about the hip."""
    rng = np.random.default_rng(seed)
    Q_ref = np.asarray(Q_ref, dtype=np.float64).reshape(-1)
    N = Q_ref.size // 12
    axis = rng.normal(size=(batch, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    R = rotation_from_rotvec(axis * rng.uniform(-angle_max, angle_max, (batch, 1)))
    om = rng.normal(0.0, omega_sigma, (batch, 3)) if omega_sigma > 0 else np.zeros((batch, 3))
    piv = np.asarray(pivot, dtype=np.float64)
    Q = np.zeros((batch, 12 * N))
    Qd = np.zeros((batch, 12 * N))
    for i in range(N):
        for k, is_point in enumerate((False, True, True, False)):
            x = Q_ref[12 * i + 3 * k : 12 * i + 3 * k + 3]
            if is_point:
                y = (R @ (x - piv)) + piv
                yd = _cross(om, y - piv)
            else:
                y = R @ x
                yd = _cross(om, y)
            Q[:, 12 * i + 3 * k : 12 * i + 3 * k + 3] = y
            Qd[:, 12 * i + 3 * k : 12 * i + 3 * k + 3] = yd
    return Q, Qd


def perturbed_inertial_parameters(table: ModelTable, batch: int, seed: int = 0, mass_rel=0.1, com_rel=0.05, inertia_rel=0.1):
    """Config 5: per-system natural inertial parameters ``(m, c, J)``: masses x U(1-mass_rel, 1+mass_rel),
    natural COM x U(1 -/+ com_rel), pseudo-inertia scaled per principal axis x U(1 -/+ inertia_rel)
    (``J -> D J D`` with ``D = diag(sqrt(s))`` keeps symmetry and definiteness)."""
    rng = np.random.default_rng(seed)
    N = table.nb_segments
    mass = table.seg_mass[None, :] * rng.uniform(1 - mass_rel, 1 + mass_rel, (batch, N))
    com = table.seg_com[None, :, :] * rng.uniform(1 - com_rel, 1 + com_rel, (batch, N, 3))
    d = np.sqrt(rng.uniform(1 - inertia_rel, 1 + inertia_rel, (batch, N, 3)))
    J = d[..., :, None] * table.seg_J[None] * d[..., None, :]
    return mass, com, J
