"""
CPU oracle for the bioNC constrained forward-dynamics hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; nothing under ``bionc_b200/`` does (the product path fails loudly without its CUDA library).

It is a plain-numpy restatement of the reference algorithm, one system at a time, operating on the
flat model table (a ``.npz`` written by ``bionc_b200.bionc_cuda.ModelTable.save`` or any object
with the same array attributes).  It deliberately keeps the reference's *algorithm*: assemble the
dense ``K``, build the dense KKT matrix ``[[G, K^T], [K, 0]]`` and LU-solve it with
``numpy.linalg.solve`` (LAPACK dgesv) -- not the Schur-complement algorithm of the CUDA kernels --
so that the two sides of every parity test are algorithmically independent.

Pinning: ``tests/test_oracle_vs_golden.py`` checks every function here against
(a) the reference's own known-answer vectors (``tests/test_joints.py``, ``tests/test_forward_dynamics.py``,
``tests/test_ultimate_examples.py`` of the reference; copied as numbers into ``tests/golden/``) and
(b) outputs of the unmodified reference run in the build container
(``tests/golden/generate_golden.py``, which imports ``/root/reference`` through ``oracle/stubs``).

Each function cites the reference lines it restates (paths relative to the reference root).
"""

from __future__ import annotations

import numpy as np

BLK_LIN3, BLK_DOT, BLK_CLEN, BLK_SOP = 0, 1, 2, 3
BLK_ROWS = {BLK_LIN3: 3, BLK_DOT: 1, BLK_CLEN: 1, BLK_SOP: 1}
FEXT_GLOBAL_ON_PROXIMAL, FEXT_GLOBAL_LOCAL_POINT, FEXT_GLOBAL, FEXT_LOCAL, FEXT_JOINT_REACTION = 0, 1, 2, 3, 4

_ARRAYS = (
    "seg_phi_const seg_mass seg_com seg_J seg_rigid_row blk_type blk_parent blk_child "
    "blk_joint blk_row_h blk_row_j blk_param"
).split()


class OracleModel:
    """Numpy model evaluated exactly like ``bionc_numpy.BiomechanicalModel`` (one system per call)."""

    def __init__(self, table, seg_mass=None, seg_com=None, seg_J=None):
        if isinstance(table, (str, bytes)) or hasattr(table, "__fspath__"):
            with np.load(table, allow_pickle=False) as z:
                d = {k: np.array(z[k]) for k in _ARRAYS}
        else:
            d = {k: np.array(getattr(table, k)) for k in _ARRAYS}
        self.__dict__.update(d)
        # markers are optional (model table format 2)
        if isinstance(table, (str, bytes)) or hasattr(table, "__fspath__"):
            with np.load(table, allow_pickle=False) as z:
                has = "marker_segment" in z.files
                self.marker_segment = np.array(z["marker_segment"]) if has else np.zeros(0, dtype=np.int32)
                self.marker_position = np.array(z["marker_position"]) if has else np.zeros((0, 3))
        else:
            ms = getattr(table, "marker_segment", None)
            self.marker_segment = np.zeros(0, dtype=np.int32) if ms is None else np.array(ms)
            mp = getattr(table, "marker_position", None)
            self.marker_position = np.zeros((0, 3)) if mp is None else np.array(mp).reshape(-1, 3)
        # optional per-system inertial-parameter override (config 5: perturbed inertial parameters)
        if seg_mass is not None:
            self.seg_mass = np.asarray(seg_mass, dtype=np.float64)
        if seg_com is not None:
            self.seg_com = np.asarray(seg_com, dtype=np.float64)
        if seg_J is not None:
            self.seg_J = np.asarray(seg_J, dtype=np.float64)
        self.nb_segments = int(self.seg_mass.shape[0])
        self.nb_Q = 12 * self.nb_segments
        self.nb_blocks = int(self.blk_type.shape[0])
        self.nb_joint_constraints = int(sum(BLK_ROWS[int(t)] for t in self.blk_type))
        self.nb_holonomic_constraints = 6 * self.nb_segments + self.nb_joint_constraints
        self._G = self._mass_matrix()
        self._fg = self._gravity_forces()

    # ----------------------------------------------------------------- constants
    def _mass_matrix(self):
        """natural_inertial_parameters.py:136-179 (G_i) and biomechanical_model.py:77-96 (block diagonal)."""
        n = self.nb_Q
        G = np.zeros((n, n))
        eye3 = np.eye(3)
        for i in range(self.nb_segments):
            m, c, J = self.seg_mass[i], self.seg_com[i], self.seg_J[i]
            Gi = np.zeros((12, 12))
            Gi[0:3, 0:3] = J[0, 0] * eye3
            Gi[0:3, 3:6] = (m * c[0] + J[0, 1]) * eye3
            Gi[0:3, 6:9] = -J[0, 1] * eye3
            Gi[0:3, 9:12] = J[0, 2] * eye3
            Gi[3:6, 3:6] = (m + 2 * m * c[1] + J[1, 1]) * eye3
            Gi[3:6, 6:9] = -(m * c[1] + J[1, 1]) * eye3
            Gi[3:6, 9:12] = (m * c[2] + J[1, 2]) * eye3
            Gi[6:9, 6:9] = J[1, 1] * eye3
            Gi[6:9, 9:12] = -J[1, 2] * eye3
            Gi[9:12, 9:12] = J[2, 2] * eye3
            Gi = np.triu(Gi) + np.triu(Gi, 1).T
            G[12 * i : 12 * i + 12, 12 * i : 12 * i + 12] = Gi
        return G

    @property
    def mass_matrix(self):
        return self._G

    def _gravity_forces(self):
        """natural_segment.py:562-572 and biomechanical_model.py:319-333."""
        f = np.zeros(self.nb_Q)
        for i in range(self.nb_segments):
            c = self.seg_com[i]
            N_com = np.kron(np.array([[c[0], 1 + c[1], -c[1], c[2]]]), np.eye(3))  # natural_vector.py:54-61
            f[12 * i : 12 * i + 12] = (N_com.T * self.seg_mass[i]) @ np.array([0, 0, -9.81])
        return f

    def gravity_forces(self):
        return self._fg.copy()

    # ----------------------------------------------------------------- energies
    def kinetic_energy(self, Qdot):
        """biomechanical_model.py:98-113: 1/2 Qdot^T G Qdot."""
        Qdot = np.asarray(Qdot, float).reshape(-1)
        return 0.5 * Qdot @ self._G @ Qdot

    def potential_energy(self, Q):
        """biomechanical_model.py:115-133 summing natural_segment.py:775-789: sum_i m_i (N_com,i Q_i)[z]
        (as coded: no factor g)."""
        Q = np.asarray(Q, float).reshape(-1)
        E = 0.0
        for i in range(self.nb_segments):
            c = self.seg_com[i]
            E += self.seg_mass[i] * (np.kron(np.array([[c[0], 1 + c[1], -c[1], c[2]]]), np.eye(3)) @ Q[12 * i : 12 * i + 12])[2]
        return E

    # ----------------------------------------------------------------- kinematics of points (post-processing)
    def center_of_mass_position(self, Q):
        """biomechanical_model_markers.py:59-79 / natural_segment.py:551-560: N_com,i Q_i -> (3, N)."""
        Q = np.asarray(Q, float).reshape(-1)
        out = np.zeros((3, self.nb_segments))
        for i in range(self.nb_segments):
            c = self.seg_com[i]
            out[:, i] = self._N([c[0], 1 + c[1], -c[1], c[2]]) @ Q[12 * i : 12 * i + 12]
        return out

    def markers(self, Q):
        """biomechanical_model_markers.py:32-57: interpolation matrix of every marker times its segment's Q_i
        (natural_marker.py:159-172, natural_vector.py interpolate()) -> (3, nb_markers)."""
        Q = np.asarray(Q, float).reshape(-1)
        out = np.zeros((3, self.marker_segment.shape[0]))
        for k, (i, p) in enumerate(zip(self.marker_segment, self.marker_position)):
            out[:, k] = self._N([p[0], 1 + p[1], -p[1], p[2]]) @ Q[12 * int(i) : 12 * int(i) + 12]
        return out

    def total_rigid_body_constraints(self, Q):
        """time_series_utils.py:6-12, :24-26 for one frame."""
        return float(np.sqrt(np.sum(self.rigid_body_constraints(Q) ** 2)))

    def total_joint_constraints(self, Q):
        """time_series_utils.py:6-12, :34-36 for one frame."""
        return float(np.sqrt(np.sum(self.joint_constraints(Q) ** 2)))

    def segment_differential_algebraic_equation(self, i, Qi, Qdoti, stabilization=None):
        """natural_segment.py:574-633: the 18x18 system [G_i Kr^T; Kr 0][Qddot_i; lambda_i] = [gravity_i; -bias_i].
        The reference's own `stabilization` branch cannot run (:617 in-place shape error); the formula it intends
        is applied here when a dict is given."""
        Qi, Qdoti = np.asarray(Qi, float).reshape(-1), np.asarray(Qdoti, float).reshape(-1)
        G = self._G[12 * i : 12 * i + 12, 12 * i : 12 * i + 12]
        Kr = self._rigid_jacobian(Qi)
        bias = -self._rigid_bias(Qdoti)
        if stabilization is not None:
            u, v, w = self._uvw(Qi)
            c = self.seg_phi_const[i]
            phi = np.array([u @ u - 1, u @ v - c[0], u @ w - c[1], v @ v - c[2], v @ w - c[3], w @ w - 1])
            bias = bias - stabilization["alpha"] * phi - stabilization["beta"] * (Kr @ Qdoti)
        A = np.zeros((18, 18))
        A[:12, :12], A[12:, :12], A[:12, 12:] = G, Kr, Kr.T
        x = np.linalg.solve(A, np.concatenate([self._fg[12 * i : 12 * i + 12], bias]))
        return x[:12], x[12:]

    # ----------------------------------------------------------------- rigid body (per segment)
    @staticmethod
    def _uvw(Qi):
        return Qi[0:3], Qi[3:6] - Qi[6:9], Qi[9:12]  # natural_coordinates.py:51-69

    def rigid_body_constraints(self, Q):
        """natural_segment.py:429-448, stacked by biomechanical_model_segments.py:26-42."""
        Q = np.asarray(Q, float).reshape(-1)
        out = np.zeros(6 * self.nb_segments)
        for i in range(self.nb_segments):
            u, v, w = self._uvw(Q[12 * i : 12 * i + 12])
            Lcg, cb, L2, Lca = self.seg_phi_const[i]
            out[6 * i : 6 * i + 6] = (u @ u - 1, u @ v - Lcg, u @ w - cb, v @ v - L2, v @ w - Lca, w @ w - 1)
        return out

    @staticmethod
    def _rigid_jacobian(Qi):
        """natural_segment.py:450-483."""
        u, v, w = OracleModel._uvw(Qi)
        Kr = np.zeros((6, 12))
        Kr[0, 0:3] = 2 * u
        Kr[1, 0:3], Kr[1, 3:6], Kr[1, 6:9] = v, u, -u
        Kr[2, 0:3], Kr[2, 9:12] = w, u
        Kr[3, 3:6], Kr[3, 6:9] = 2 * v, -2 * v
        Kr[4, 3:6], Kr[4, 6:9], Kr[4, 9:12] = w, -w, v
        Kr[5, 9:12] = 2 * w
        return Kr

    def rigid_body_constraints_jacobian(self, Q):
        """biomechanical_model_segments.py:62-79."""
        Q = np.asarray(Q, float).reshape(-1)
        K = np.zeros((6 * self.nb_segments, self.nb_Q))
        for i in range(self.nb_segments):
            K[6 * i : 6 * i + 6, 12 * i : 12 * i + 12] = self._rigid_jacobian(Q[12 * i : 12 * i + 12])
        return K

    @staticmethod
    def _rigid_bias(Qdi):
        """natural_segment.py:508-545 with the constant Hessians of :848-923 contracted by hand."""
        ud, vd, wd = OracleModel._uvw(Qdi)
        return 2 * np.array([ud @ ud, ud @ vd, ud @ wd, vd @ vd, vd @ wd, wd @ wd])

    # ----------------------------------------------------------------- joint blocks
    @staticmethod
    def _N(coefs):
        return np.kron(np.asarray(coefs, float).reshape(1, 4), np.eye(3))

    def _block_eval(self, b, Q, Qdot):
        """Constraint, parent/child Jacobian blocks and bias of primitive block ``b``.
        LIN3: joints.py:523-587 (Spherical), :201/:380 (Hinge/Universal translation),
              joints_with_ground.py:153, :304, :401-426, :505-528
        DOT : joints.py:203-208, :222-235, :252-282 ; joints_with_ground.py:155-162, :169-178
        CLEN: joints.py:1300-1382
        SOP : joints.py:657-748 (Jacobian blocks stored exactly where the reference stores them)
        """
        t = int(self.blk_type[b])
        p, c = int(self.blk_parent[b]), int(self.blk_child[b])
        P = self.blk_param[b]
        Qc = Q[12 * c : 12 * c + 12]
        Qp = Q[12 * p : 12 * p + 12] if p >= 0 else None
        Qdc = None if Qdot is None else Qdot[12 * c : 12 * c + 12]
        Qdp = None if (Qdot is None or p < 0) else Qdot[12 * p : 12 * p + 12]
        nr = BLK_ROWS[t]
        phi = np.zeros(nr)
        Kp = np.zeros((nr, 12)) if p >= 0 else None
        Kc = np.zeros((nr, 12))
        bias = np.zeros(nr)
        if t == BLK_LIN3:
            Np, Nc, c0 = self._N(P[0:4]), self._N(P[4:8]), P[8:11]
            phi[:] = c0 - Nc @ Qc
            Kc[:] = -Nc
            if p >= 0:
                phi += Np @ Qp
                Kp[:] = Np
        elif t == BLK_DOT:
            Nc = self._N(P[4:8])
            dc = Nc @ Qc
            if p >= 0:
                Np = self._N(P[0:4])
                dp = Np @ Qp
                Kp[0] = dc @ Np
                if Qdot is not None:
                    bias[0] = 2 * (Np @ Qdp) @ (Nc @ Qdc)
            else:
                dp = P[0:3]
            phi[0] = dp @ dc - P[8]
            Kc[0] = dp @ Nc
        elif t == BLK_CLEN:
            Np, Nc = self._N(P[0:4]), self._N(P[4:8])
            d = Np @ Qp - Nc @ Qc
            phi[0] = np.sum(d**2) - P[8]
            Kp[0] = 2 * d @ Np
            Kc[0] = -2 * d @ Nc
            if Qdot is not None:
                dd = Np @ Qdp - Nc @ Qdc
                bias[0] = 2 * dd @ dd
        elif t == BLK_SOP:
            NP, NA, Nn, r = self._N(P[0:4]), self._N(P[4:8]), self._N(P[8:12]), P[12]
            Pp, A, n = NP @ Qp, NA @ Qc, Nn @ Qc
            phi[0] = (Pp - A) @ n - r
            Kp[0] = -n @ NA + (Pp - A) @ Nn  # joints.py:676-690: stored in the PARENT columns as coded
            Kc[0] = n @ NP  # joints.py:692-698: stored in the CHILD columns as coded
            if Qdot is not None:
                Pd, Ad, nd = NP @ Qdp, NA @ Qdc, Nn @ Qdc
                bias[0] = 2 * Pd @ nd - 2 * nd @ Ad
        else:
            raise ValueError(f"unknown block type {t}")
        return phi, Kp, Kc, bias

    def _row_offsets(self, order):
        return self.blk_row_h if order == "holonomic" else self.blk_row_j

    def joint_constraints(self, Q):
        """biomechanical_model_joints.py:13-41 (joint-index order)."""
        Q = np.asarray(Q, float).reshape(-1)
        out = np.zeros(self.nb_joint_constraints)
        for b in range(self.nb_blocks):
            phi, _, _, _ = self._block_eval(b, Q, None)
            out[self.blk_row_j[b] : self.blk_row_j[b] + phi.size] = phi
        return out

    def joint_constraints_jacobian(self, Q):
        """biomechanical_model_joints.py:43-77 (joint-index order)."""
        Q = np.asarray(Q, float).reshape(-1)
        K = np.zeros((self.nb_joint_constraints, self.nb_Q))
        for b in range(self.nb_blocks):
            _, Kp, Kc, _ = self._block_eval(b, Q, None)
            r0, p, c = self.blk_row_j[b], int(self.blk_parent[b]), int(self.blk_child[b])
            if p >= 0:
                K[r0 : r0 + Kc.shape[0], 12 * p : 12 * p + 12] = Kp
            K[r0 : r0 + Kc.shape[0], 12 * c : 12 * c + 12] = Kc
        return K

    def holonomic_constraints(self, Q):
        """biomechanical_model.py:154-197."""
        Q = np.asarray(Q, float).reshape(-1)
        phi = np.zeros(self.nb_holonomic_constraints)
        phir = self.rigid_body_constraints(Q)
        for b in range(self.nb_blocks):
            ph, _, _, _ = self._block_eval(b, Q, None)
            phi[self.blk_row_h[b] : self.blk_row_h[b] + ph.size] = ph
        for i in range(self.nb_segments):
            phi[self.seg_rigid_row[i] : self.seg_rigid_row[i] + 6] = phir[6 * i : 6 * i + 6]
        return phi

    def holonomic_constraints_jacobian(self, Q):
        """biomechanical_model.py:199-258."""
        Q = np.asarray(Q, float).reshape(-1)
        K = np.zeros((self.nb_holonomic_constraints, self.nb_Q))
        for b in range(self.nb_blocks):
            _, Kp, Kc, _ = self._block_eval(b, Q, None)
            r0, p, c = self.blk_row_h[b], int(self.blk_parent[b]), int(self.blk_child[b])
            K[r0 : r0 + Kc.shape[0], 12 * c : 12 * c + 12] = Kc
            if p >= 0:
                K[r0 : r0 + Kc.shape[0], 12 * p : 12 * p + 12] = Kp
        for i in range(self.nb_segments):
            r0 = self.seg_rigid_row[i]
            K[r0 : r0 + 6, 12 * i : 12 * i + 12] = self._rigid_jacobian(Q[12 * i : 12 * i + 12])
        return K

    def holonomic_constraints_acceleration_bias(self, Qdot):
        """biomechanical_model.py:260-317."""
        Qdot = np.asarray(Qdot, float).reshape(-1)
        bias = np.zeros(self.nb_holonomic_constraints)
        # joint biases depend on Qdot only; any Q of the right size serves the shared evaluator
        for b in range(self.nb_blocks):
            _, _, _, bb = self._block_eval(b, Qdot, Qdot)
            bias[self.blk_row_h[b] : self.blk_row_h[b] + bb.size] = bb
        for i in range(self.nb_segments):
            r0 = self.seg_rigid_row[i]
            bias[r0 : r0 + 6] = self._rigid_bias(Qdot[12 * i : 12 * i + 12])
        return bias

    # ----------------------------------------------------------------- external forces
    def natural_external_forces(self, Q, fext):
        """external_force.py:143-180.  ``fext`` is a list of ``(segment, kind, param24)`` with
        param = [torque(3), force(3), point(3), Tinv(9)] (torque first: external_force.py:84-85)."""
        Q = np.asarray(Q, float).reshape(-1)
        f = np.zeros(self.nb_Q)
        if not fext:
            return f
        for seg, kind, par in fext:
            par = np.asarray(par, float)
            Qi = Q[12 * seg : 12 * seg + 12]
            u, rp, rd, w = Qi[0:3], Qi[3:6], Qi[6:9], Qi[9:12]
            v = rp - rd
            tau, F, pt = par[0:3].copy(), par[3:6].copy(), par[6:9]
            if kind == FEXT_LOCAL:  # external_force_in_local.py:85-131
                R = np.column_stack((u, v, w)) @ par[9:18].reshape(3, 3)
                F, tau = R @ F, R @ tau
            if kind in (FEXT_GLOBAL_LOCAL_POINT, FEXT_LOCAL):  # external_force_global_local_point.py:68-100
                P_app = pt[0] * u + (1 + pt[1]) * rp - pt[1] * rd + pt[2] * w
                tau = tau + np.cross(rp - P_app, F)
            elif kind == FEXT_GLOBAL:  # external_force_global.py:67-96
                tau = tau + np.cross(rp - pt, F)
            elif kind == FEXT_JOINT_REACTION:
                # reaction of a joint generalized force on the joint's parent (generalized_force.py:98-129): the force acts on
                # the child par[18] at its proximal point; transported to this segment's proximal point as
                # external_force_global_on_proximal.py:117-147 codes it (lever = new point - old point), then negated
                rp_child = Q[12 * int(par[18]) + 3 : 12 * int(par[18]) + 6]
                tau = -(tau + np.cross(rp - rp_child, F))
                F = -F
            # external_force_global_on_proximal.py:65-115 + natural_coordinates.py:132-158
            left = np.zeros((12, 3))
            left[9:12, 0] = u
            left[0:3, 1] = v
            left[3:6, 2] = -w
            left[6:9, 2] = w
            Bstar = np.column_stack((np.cross(w, u), np.cross(u, v), np.cross(-v, w)))
            fi = left @ np.linalg.inv(Bstar) @ tau
            fi[3:6] += F
            f[12 * seg : 12 * seg + 12] += fi
        return f

    # ----------------------------------------------------------------- dynamics
    def augmented_mass_matrix(self, Q):
        """biomechanical_model.py:335-349."""
        K = self.holonomic_constraints_jacobian(Q)
        m = K.shape[0]
        return np.block([[self._G, K.T], [K, np.zeros((m, m))]])

    def forward_dynamics(self, Q, Qdot, fext=None, stabilization=None):
        """biomechanical_model.py:351-429: dense KKT, LAPACK LU solve."""
        Q = np.asarray(Q, float).reshape(-1)
        Qdot = np.asarray(Qdot, float).reshape(-1)
        A = self.augmented_mass_matrix(Q)
        forces = self._fg + self.natural_external_forces(Q, fext)
        bias = -self.holonomic_constraints_acceleration_bias(Qdot)
        if stabilization is not None:
            bias = bias - (
                stabilization["alpha"] * self.holonomic_constraints(Q)
                + stabilization["beta"] * self.holonomic_constraints_jacobian(Q) @ Qdot
            )
        x = np.linalg.solve(A, np.concatenate([forces, bias]))
        return x[: self.nb_Q], x[self.nb_Q :]

    def rk4(self, Q0, Qdot0, t, normalize=True, fext=None, stabilization=None, return_all=False):
        """utils/ode_solver.py:8-53 with the closure of :107-116; ``t`` is the time grid
        (``h = t[i+1] - t[i]`` recomputed every step, :41).  ``normalize`` = passing
        ``model.normalized_coordinates`` (protocols/biomechanical_model_segments.py:226-236)."""
        n = self.nb_Q
        y = np.concatenate([np.asarray(Q0, float).reshape(-1), np.asarray(Qdot0, float).reshape(-1)])
        hist = [y.copy()] if return_all else None

        def f(yy):
            qdd, _ = self.forward_dynamics(yy[:n], yy[n:], fext, stabilization)
            return np.concatenate([yy[n:], qdd])

        for i in range(len(t) - 1):
            h = t[i + 1] - t[i]
            k1 = f(y)
            k2 = f(y + k1 * h / 2.0)
            k3 = f(y + k2 * h / 2.0)
            k4 = f(y + k3 * h)
            y = y + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            if normalize:
                for s in range(self.nb_segments):
                    for sl in (slice(12 * s, 12 * s + 3), slice(12 * s + 9, 12 * s + 12)):
                        y[sl] = y[sl] / np.linalg.norm(y[sl])
            if return_all:
                hist.append(y.copy())
        if return_all:
            return np.array(hist).T
        return y[:n], y[n:]


def rk4_generic(t, f, y0, normalize_idx=None, args=()):
    """utils/ode_solver.py:8-53 for an arbitrary right-hand side (pinned by the reference's
    tests/test_ode_solver.py:6-53)."""
    n = len(t)
    y = np.zeros((len(y0), n))
    y[:, 0] = y0
    for i in range(n - 1):
        h = t[i + 1] - t[i]
        yi = np.squeeze(y[:, i])
        k1 = f(t[i], yi, *args)
        k2 = f(t[i] + h / 2.0, yi + k1 * h / 2.0, *args)
        k3 = f(t[i] + h / 2.0, yi + k2 * h / 2.0, *args)
        k4 = f(t[i] + h, yi + k3 * h, *args)
        y[:, i + 1] = yi + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        if normalize_idx is not None:
            for idx in normalize_idx:
                y[idx, i + 1] = y[idx, i + 1] / np.linalg.norm(y[idx, i + 1])
    return y
