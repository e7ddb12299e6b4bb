"""``bionc_cuda.BiomechanicalModel``: the batched CUDA mirror of
``bionc.bionc_numpy.BiomechanicalModel`` for the constrained forward-dynamics path.

Same method names and argument meaning as the reference
(``bionc/bionc_numpy/biomechanical_model.py``, ``bionc/protocols/biomechanical_model.py:347-481``);
every state argument gains a leading batch axis:

===============================  =======================  ==========================
reference (one system)           bionc_cuda (single)      bionc_cuda (batch)
===============================  =======================  ==========================
``Q``, ``Qdot``: ``(n, 1)``      ``(n,)`` or ``(n, 1)``   ``(B, n)``
``Qddot``: ``(n, 1)``            ``(n, 1)``               ``(B, n)``
``lambdas``: ``(m, 1)``          ``(m, 1)``               ``(B, m)``
``K``: ``(m, n)``                ``(m, n)``               ``(B, m, n)``
===============================  =======================  ==========================

numpy in -> numpy out; torch CUDA tensors in -> torch CUDA tensors out (no host round trip).
All arithmetic happens in the hand-written sm_100a kernels of ``bionc_b200/csrc`` reached through
the C ABI of ``include/bionc_cuda.h``; there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .external_force import ExternalForceSet
from .model_table import ModelTable

EVAL_PHI_R, EVAL_K_R, EVAL_PHI_J, EVAL_K_J, EVAL_PHI_H, EVAL_K_H, EVAL_BIAS_H, EVAL_FEXT = range(8)
EVAL_COM, EVAL_MARKERS, EVAL_KINETIC, EVAL_POTENTIAL, EVAL_RESIDUALS = range(8, 13)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class _Batch:
    """Normalises user input to a contiguous float64 CUDA tensor (B, n) and remembers how to go back."""

    def __init__(self, x, n: int, device: torch.device, name: str):
        self.torch_in = isinstance(x, torch.Tensor)
        if self.torch_in:
            t = x
            shape = tuple(t.shape)
        else:
            a = np.asarray(x, dtype=np.float64)
            shape = a.shape
            t = torch.from_numpy(np.ascontiguousarray(a))
        self.single = len(shape) == 1 or (len(shape) == 2 and shape[1] == 1 and shape[0] == n)
        if self.single:
            t = t.reshape(1, -1)
        if t.dim() != 2 or t.shape[1] != n:
            raise ValueError(f"{name} must have shape ({n},), ({n}, 1) or (B, {n}); got {shape}")
        self.t = t.to(device=device, dtype=torch.float64).contiguous()
        if self.t.data_ptr() % 16:  # a view at an odd double of a larger buffer: the kernels use 16-byte accesses
            self.t = self.t.clone()
        self.B = int(self.t.shape[0])

    def out(self, t: torch.Tensor, column: bool = False):
        """Back to the caller's world: torch stays on the device; numpy goes to the host; a single
        system loses the batch axis (``column`` -> trailing axis of 1 like the reference)."""
        if self.single:
            t = t[0]
            if column:
                t = t.unsqueeze(-1)
        return t if self.torch_in else t.cpu().numpy()


class BiomechanicalModel:
    """Compiled model on one CUDA device.  Build it with :meth:`from_numpy` (from a loaded
    ``bionc_numpy`` model -- the analogue of ``to_mx()``), :meth:`from_table` or :meth:`load`."""

    def __init__(self, table: ModelTable, device=None):
        self.table = table.validate()
        if not torch.cuda.is_available():
            raise _lib.BioncCudaError("bionc_cuda needs a CUDA device (torch.cuda.is_available() is False); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("bionc_cuda models live on CUDA devices only")
        self._lib = _lib.load()
        desc, self._keep = _lib.model_desc(self.table)
        handle = C.c_void_p()
        _lib.check(self._lib.bionc_cuda_model_create(C.byref(desc), self.device.index or 0, C.byref(handle)), "model_create")
        self._no_jit = False  # set when a model-specialised kernel could not be built here (rk4(specialize="auto"))
        self._h = handle
        self._mass_matrix = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self, "_lib", None) is not None:
            try:
                self._lib.bionc_cuda_model_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_numpy(cls, model, device=None) -> "BiomechanicalModel":
        """From a loaded ``bionc.bionc_numpy.BiomechanicalModel`` (e.g. ``BiomechanicalModel.load("x.nmod")``)."""
        return cls(ModelTable.from_numpy(model), device)

    @classmethod
    def from_table(cls, table: ModelTable, device=None) -> "BiomechanicalModel":
        return cls(table, device)

    @classmethod
    def load(cls, filename: str, device=None) -> "BiomechanicalModel":
        """Load a compiled model table (``.npz`` written by :meth:`save`)."""
        return cls(ModelTable.load(filename), device)

    def save(self, filename: str):
        self.table.save(filename)

    # ------------------------------------------------------------------ counts (protocols/biomechanical_model.py:347-453)
    @property
    def nb_segments(self) -> int:
        return self.table.nb_segments

    @property
    def nb_Q(self) -> int:
        return 12 * self.nb_segments

    nb_Qdot = nb_Q
    nb_Qddot = nb_Q

    @property
    def nb_rigid_body_constraints(self) -> int:
        return 6 * self.nb_segments

    @property
    def nb_joint_constraints(self) -> int:
        return self.table.nb_joint_constraints

    @property
    def nb_holonomic_constraints(self) -> int:
        return self.table.nb_holonomic_constraints

    @property
    def nb_joints(self) -> int:
        return len(self.table.joint_names)

    @property
    def segment_names(self) -> list[str]:
        return list(self.table.segment_names)

    @property
    def joint_names(self) -> list[str]:
        return list(self.table.joint_names)

    @property
    def normalized_coordinates(self) -> list[list[int]]:
        """protocols/biomechanical_model_segments.py:226-236"""
        idx = []
        for i in range(self.nb_segments):
            idx.append([12 * i, 12 * i + 1, 12 * i + 2])
            idx.append([12 * i + 9, 12 * i + 10, 12 * i + 11])
        return idx

    def external_force_set(self) -> ExternalForceSet:
        return ExternalForceSet.empty_from_nb_segment(self.nb_segments)

    # ------------------------------------------------------------------ constants
    @property
    def mass_matrix(self) -> np.ndarray:
        """Block-diagonal ``G`` (biomechanical_model.py:77-96), assembled by the library from the uploaded table."""
        if self._mass_matrix is None:
            n = self.nb_Q
            G = np.zeros((n, n))
            _lib.check(self._lib.bionc_cuda_mass_matrix(self._h, G.ctypes.data), "mass_matrix")
            self._mass_matrix = G
        return self._mass_matrix

    def gravity_forces(self) -> np.ndarray:
        """biomechanical_model.py:319-333 -> (n, 1)"""
        return self.table.seg_gravity.reshape(-1, 1)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _options(self, external_forces=None, stabilization=None, inertia=None, B=None, force_values=None, host=False,
                 joint_forces=None):
        """``BioncDynOptions`` of one call plus the objects it points into.  ``force_values``: optional per-system
        (torque, force) values ``(B, F, 6)`` -- or ``(steps, B, F, 6)`` for ``rk4`` -- replacing the values stored in
        ``external_forces`` (forces in the order ``ExternalForceSet.to_arrays`` lists them: by segment, then insertion);
        ``host``: the call goes to a ``_host`` entry point (values stay on the host)."""
        keep = []
        opt = _lib.BioncDynOptions()
        opt.use_stabilization = 0
        if stabilization is not None:
            opt.use_stabilization, opt.alpha, opt.beta = 1, float(stabilization["alpha"]), float(stabilization["beta"])
        if external_forces is not None:
            if not isinstance(external_forces, ExternalForceSet):
                external_forces = ExternalForceSet.from_reference(external_forces)
            if len(external_forces) != self.nb_segments:
                # same check and message as external_force.py:154-157
                raise ValueError(
                    "The number of segment in the model and the number of segment in the external forces must be the same"
                )
            seg, kind, par = external_forces.to_arrays()
        if joint_forces is not None:  # appended behind the external forces (apply_joint_generalized_forces=True)
            from .generalized_force import JointGeneralizedForcesList

            if not isinstance(joint_forces, JointGeneralizedForcesList):
                if not hasattr(joint_forces, "joint_generalized_forces"):
                    raise NotImplementedError(
                        "apply_joint_generalized_forces needs a JointGeneralizedForcesList: the reference's conversion of a "
                        "joint-dof array (add_all_joint_generalized_forces) is not lowered")
                joint_forces = JointGeneralizedForcesList.from_reference(joint_forces)
            if force_values is not None:
                raise ValueError("per-system external force values cannot be combined with joint generalized forces")
            if not len(self.table.all_joint_names):
                raise ValueError("this model table carries no joint list (built before format 3): rebuild it with from_numpy")
            js, jk, jp = joint_forces.to_arrays(self.table.all_joint_parent, self.table.all_joint_child)
            if external_forces is None:
                seg, kind, par = js, jk, jp
            else:
                seg, kind, par = np.concatenate([seg, js]), np.concatenate([kind, jk]), np.concatenate([par, jp])
                seg, kind, par = np.ascontiguousarray(seg), np.ascontiguousarray(kind), np.ascontiguousarray(par)
            external_forces = True
        if external_forces is not None:
            if seg.size:
                fx = _lib.BioncFextDesc(int(seg.size), _ptr(seg, _lib.c_int32_p), _ptr(kind, _lib.c_int32_p), _ptr(par, _lib.c_double_p))
                if force_values is not None:
                    F = int(seg.size)
                    if host:
                        fv = np.ascontiguousarray(force_values, dtype=np.float64)
                    else:
                        fv = force_values if isinstance(force_values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(force_values, dtype=np.float64))
                        fv = fv.to(device=self.device, dtype=torch.float64).contiguous()
                    shape = tuple(fv.shape)
                    if shape == (B, F, 6):
                        steps = 1
                    elif len(shape) == 4 and shape[1:] == (B, F, 6) and shape[0] >= 1:
                        steps = int(shape[0])
                    else:
                        raise ValueError(f"external force values must have shape ({B}, {F}, 6) or (steps, {B}, {F}, 6); got {shape}")
                    fx.values = fv.ctypes.data if host else fv.data_ptr()
                    fx.values_steps = steps
                    keep.append(fv)
                keep += [seg, kind, par, fx]
                opt.fext = C.pointer(fx)
            elif force_values is not None:
                raise ValueError("external force values given for an empty force set")
        elif force_values is not None:
            raise ValueError("external_force_values needs external_forces (the list of application points)")
        if inertia is not None:
            it = inertia if isinstance(inertia, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(inertia, dtype=np.float64))
            it = it.to(device=self.device, dtype=torch.float64).contiguous()
            if tuple(it.shape) != (B, self.nb_segments, 10):
                raise ValueError(f"inertia must have shape (B, {self.nb_segments}, 10) = [m, c(3), J00 J01 J02 J11 J12 J22]")
            keep.append(it)
            opt.inertia = it.data_ptr()
        keep.append(opt)
        return opt, keep

    def specialize(self, external_forces=None, stabilization=None, per_system_inertia: bool = False, joint_forces=None) -> float:
        """Compile this model's own build of the fused RK4 kernel (NVRTC, sm_100a) for calls shaped like the arguments
        -- Baumgarte on / off, external forces present or not, per-system inertial parameters or not -- and load it:
        ``rk4`` / ``rk4_host`` calls of that shape then run it instead of the generic kernel (the model tables are
        compile-time constants in it; same results to rounding).  The counterpart, for this
        path, of the reference's code generation ``model.to_mx()`` (bionc_casadi/biomechanical_model.py).  Returns the
        host seconds spent (0.0 if that shape was compiled already).

        Measured: it pays for models with other than point-to-point joints and for calls with
        forces, up to about four segments (pendulum 2.8x, knee 3.3x, 4-link hinge chain 1.5x); models on the lean path
        (point-to-point joints only, no forces: lower limb, full body) are faster on the generic kernel, whose code all
        warps share.  ``rk4(specialize="auto")`` applies exactly that rule."""
        import time

        opt, keep = self._options(external_forces, stabilization, None, None, None, joint_forces=joint_forces)
        if per_system_inertia:
            opt.inertia = 1  # only the presence is read
        before = self._lib.bionc_cuda_model_specialized(self._h, None)
        t0 = time.perf_counter()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bionc_cuda_model_specialize(self._h, C.byref(opt), 0, 1), "model_specialize")
        dt = time.perf_counter() - t0
        return dt if self._lib.bionc_cuda_model_specialized(self._h, None) > before else 0.0

    def specialized_kernels(self) -> int:
        """number of loaded model-specialised kernels"""
        return int(self._lib.bionc_cuda_model_specialized(self._h, None))

    def _eval(self, which: int, x, rows: int | None, cols: int | None, name: str, external_forces=None, force_values=None):
        b = _Batch(x, self.nb_Q, self.device, name)
        shape = (b.B,) + ((rows,) if rows is not None else ()) + ((cols,) if cols is not None else ())
        out = torch.empty(shape, dtype=torch.float64, device=self.device)
        opt, keep = self._options(external_forces, B=b.B, force_values=force_values)
        with torch.cuda.device(self.device):
            if out.numel():
                _lib.check(self._lib.bionc_cuda_eval(self._h, which, b.t.data_ptr(), C.byref(opt), out.data_ptr(), b.B, self._stream()), "eval")
        return b, out

    # ------------------------------------------------------------------ constraint API
    def rigid_body_constraints(self, Q):
        """biomechanical_model_segments.py:26-42 -> (6N,) / (B, 6N)"""
        b, out = self._eval(EVAL_PHI_R, Q, 6 * self.nb_segments, None, "Q")
        return b.out(out)

    def rigid_body_constraints_jacobian(self, Q):
        """biomechanical_model_segments.py:62-79 -> (6N, n) / (B, 6N, n)"""
        b, out = self._eval(EVAL_K_R, Q, 6 * self.nb_segments, self.nb_Q, "Q")
        return b.out(out)

    def joint_constraints(self, Q):
        """biomechanical_model_joints.py:13-41 (joint-index order) -> (mj,) / (B, mj)"""
        b, out = self._eval(EVAL_PHI_J, Q, self.nb_joint_constraints, None, "Q")
        return b.out(out)

    def joint_constraints_jacobian(self, Q):
        """biomechanical_model_joints.py:43-77 -> (mj, n) / (B, mj, n)"""
        b, out = self._eval(EVAL_K_J, Q, self.nb_joint_constraints, self.nb_Q, "Q")
        return b.out(out)

    def holonomic_constraints(self, Q):
        """biomechanical_model.py:154-197 -> (m, 1) / (B, m)"""
        b, out = self._eval(EVAL_PHI_H, Q, self.nb_holonomic_constraints, None, "Q")
        return b.out(out, column=True)

    def holonomic_constraints_jacobian(self, Q):
        """biomechanical_model.py:199-258 -> (m, n) / (B, m, n)"""
        b, out = self._eval(EVAL_K_H, Q, self.nb_holonomic_constraints, self.nb_Q, "Q")
        return b.out(out)

    def holonomic_constraints_acceleration_bias(self, Qdot):
        """biomechanical_model.py:260-317 -> (m, 1) / (B, m)"""
        b, out = self._eval(EVAL_BIAS_H, Qdot, self.nb_holonomic_constraints, None, "Qdot")
        return b.out(out, column=True)

    def natural_external_forces(self, Q, external_forces, external_force_values=None):
        """ExternalForceSet.to_natural_external_forces (external_force.py:143-166) -> (n, 1) / (B, n);
        ``external_force_values`` (B, F, 6): per-system (torque, force) values."""
        b, out = self._eval(EVAL_FEXT, Q, self.nb_Q, None, "Q", external_forces=external_forces, force_values=external_force_values)
        return b.out(out, column=True)

    def augmented_mass_matrix(self, Q):
        """biomechanical_model.py:335-349 -> (n+m, n+m) / (B, n+m, n+m); assembled with torch from K (not on the hot path)."""
        b, K = self._eval(EVAL_K_H, Q, self.nb_holonomic_constraints, self.nb_Q, "Q")
        n, m = self.nb_Q, self.nb_holonomic_constraints
        A = torch.zeros((b.B, n + m, n + m), dtype=torch.float64, device=self.device)
        A[:, :n, :n] = torch.from_numpy(self.mass_matrix).to(self.device)
        A[:, n:, :n] = K
        A[:, :n, n:] = K.transpose(1, 2)
        return b.out(A)

    # ------------------------------------------------------------------ post-processing on the device
    def kinetic_energy(self, Qdot):
        """biomechanical_model.py:98-113: ``1/2 Qdot^T G Qdot`` -> float / (B,)"""
        b, E = self._eval(EVAL_KINETIC, Qdot, None, None, "Qdot")
        return float(E[0].item()) if (b.single and not b.torch_in) else b.out(E)

    def potential_energy(self, Q):
        """biomechanical_model.py:115-133: ``sum_i m_i (N_com,i Q_i)[z]`` exactly as the reference codes it
        (``natural_segment.py:775-789``, no factor g) -> float / (B,)"""
        b, E = self._eval(EVAL_POTENTIAL, Q, None, None, "Q")
        return float(E[0].item()) if (b.single and not b.torch_in) else b.out(E)

    @property
    def nb_markers(self) -> int:
        return self.table.nb_markers

    @property
    def nb_markers_technical(self) -> int:
        return int(self.table.marker_technical.sum())

    @property
    def marker_names(self) -> list[str]:
        return list(self.table.marker_names)

    @property
    def marker_names_technical(self) -> list[str]:
        return [n for n, t in zip(self.table.marker_names, self.table.marker_technical) if t]

    def markers(self, Q):
        """Forward kinematics, biomechanical_model_markers.py:32-57 -> (3, nb_markers, 1) for one system like the
        reference, (B, 3, nb_markers) for a batch."""
        b, out = self._eval(EVAL_MARKERS, Q, 3, self.nb_markers, "Q")
        return self._points_out(b, out)

    def center_of_mass_position(self, Q):
        """biomechanical_model_markers.py:59-79 -> (3, N, 1) for one system, (B, 3, N) for a batch."""
        b, out = self._eval(EVAL_COM, Q, 3, self.nb_segments, "Q")
        return self._points_out(b, out)

    @staticmethod
    def _points_out(b, out):
        if b.single:
            out = out[0].unsqueeze(-1)  # the reference's trailing frame axis
            return out if b.torch_in else out.cpu().numpy()
        return out if b.torch_in else out.cpu().numpy()

    def constraint_residual_norms(self, Q):
        """``[|Phi_r(Q)|_2, |Phi_j(Q)|_2]`` per system: the two numbers ``time_series_utils.total_rigid_body_constraints``
        / ``total_joint_constraints`` (bionc_numpy/time_series_utils.py:6-41) compute per frame -> (2,) / (B, 2)."""
        b, out = self._eval(EVAL_RESIDUALS, Q, 2, None, "Q")
        return b.out(out)

    def segment(self, index):
        """Single-segment view (``model.segments[name]`` of the reference), by index or name."""
        from .natural_segment import NaturalSegment

        if isinstance(index, str):
            index = self.table.segment_names.index(index)
        return NaturalSegment(self, int(index))

    def lagrangian(self, Q, Qdot):
        """biomechanical_model.py:135-152"""
        return self.kinetic_energy(Qdot) - self.potential_energy(Q)

    # ------------------------------------------------------------------ dynamics
    def forward_dynamics(
        self,
        Q,
        Qdot,
        joint_generalized_forces=None,
        external_forces=None,
        stabilization: dict = None,
        inertial_parameters=None,
        return_status: bool = False,
        pivoting: str = "auto",
        external_force_values=None,
        apply_joint_generalized_forces: bool = False,
    ):
        """biomechanical_model.py:351-429, batched.  ``joint_generalized_forces`` is accepted and ignored
        exactly like the reference (:391-413) unless ``apply_joint_generalized_forces=True``: then a
        ``JointGeneralizedForcesList`` (``bionc_cuda.generalized_force``) acts on the joints' children and reacts on
        their parents as generalized_force.py:98-129 documents.  ``inertial_parameters`` (B, N, 10) optionally overrides
        the natural inertial parameters per system; ``external_force_values`` (B, F, 6) gives every system its own
        (torque, force) for the forces of ``external_forces`` (the reference builds one ``ExternalForceSet`` per call,
        external_force.py:143-180).  Returns ``(Qddot, lambdas)``; a single system whose
        matrix is singular raises ``numpy.linalg.LinAlgError`` like the reference's solve.

        ``pivoting``: ``"auto"`` (default) re-solves the systems the fast block-elimination kernel flags
        (zero / small / non-finite pivot of the unpivoted LDL^T, e.g. indefinite mass matrices) with the dense
        pivoted-LU kernel -- the reference's own algorithm (:335-349, :423-426) on the device -- and marks them
        ``BIONC_STATUS_PIVOTED``; ``"never"`` keeps the fast result and its flags (no host synchronisation);
        ``"always"`` solves every system with the dense kernel."""
        if pivoting not in ("auto", "never", "always"):
            raise ValueError("pivoting must be 'auto', 'never' or 'always'")
        q = _Batch(Q, self.nb_Q, self.device, "Q")
        qd = _Batch(Qdot, self.nb_Q, self.device, "Qdot")
        if q.B != qd.B:
            raise ValueError("Q and Qdot must have the same batch size")
        B = q.B
        qdd = torch.empty_like(q.t)
        lam = torch.empty((B, self.nb_holonomic_constraints), dtype=torch.float64, device=self.device)
        status = torch.empty((B,), dtype=torch.int32, device=self.device)
        opt, keep = self._options(external_forces, stabilization, inertial_parameters, B, external_force_values,
                                  joint_forces=joint_generalized_forces if apply_joint_generalized_forces else None)
        with torch.cuda.device(self.device):
            if B > 0 and pivoting != "always":
                rc = self._lib.bionc_cuda_forward_dynamics(
                    self._h, q.t.data_ptr(), qd.t.data_ptr(), C.byref(opt), qdd.data_ptr(), lam.data_ptr(),
                    status.data_ptr(), B, self._stream(),
                )
                _lib.check(rc, "forward_dynamics")
            if B > 0 and pivoting != "never":
                select = None if pivoting == "always" else torch.nonzero(status & 7).flatten()
                count = B if select is None else int(select.numel())
                if count > 0:
                    nbytes = int(self._lib.bionc_cuda_dense_workspace_bytes(self._h, count))
                    work = torch.empty((nbytes // 8,), dtype=torch.float64, device=self.device)
                    rc = self._lib.bionc_cuda_forward_dynamics_dense(
                        self._h, q.t.data_ptr(), qd.t.data_ptr(), C.byref(opt), None if select is None else select.data_ptr(),
                        count, qdd.data_ptr(), lam.data_ptr(), status.data_ptr(), work.data_ptr(), nbytes, self._stream(),
                    )
                    _lib.check(rc, "forward_dynamics_dense")
        if q.single and B > 0 and not return_status:
            if int(status[0].item()) & 1:
                raise np.linalg.LinAlgError("Singular matrix")
        res = (q.out(qdd, column=True), q.out(lam, column=True))
        return res + (q.out(status),) if return_status else res

    def rk4(
        self,
        Q,
        Qdot,
        h: float | None = None,
        n_steps: int | None = None,
        t=None,
        normalize: bool = True,
        external_forces=None,
        stabilization: dict = None,
        inertial_parameters=None,
        save_every: int | None = None,
        return_lambdas: bool = False,
        return_status: bool = False,
        inplace: bool = False,
        pivoting: str = "never",
        external_force_values=None,
        joint_generalized_forces=None,
        apply_joint_generalized_forces: bool = False,
        specialize="auto",
    ):
        """Fused integrator: ``utils/ode_solver.py:8-53`` with the closure of ``:107-116`` inside the kernel.

        ``specialize``: ``"auto"`` compiles this model's own build of the kernel (:meth:`specialize`) the first time a
        large call (>= 2e6 system-steps) of a shape that profits from it arrives -- models with other than
        point-to-point joints, or calls with forces, of at most 4 segments; ``True`` always, ``False`` never.  Where
        NVRTC is missing, ``"auto"`` keeps the generic kernel.

        Either a constant step ``h`` with ``n_steps``, or a time grid ``t`` (then ``h_i = t[i+1] - t[i]``
        exactly as the reference computes it, :41).  ``normalize=True`` is the reference's
        ``normalize_idx=model.normalized_coordinates``.  Returns ``(Q, Qdot)`` after the last step, plus
        the trajectory ``(n_steps // save_every, B, 2n)`` if ``save_every`` is given, the multipliers of
        the first stage of the last step if ``return_lambdas``, and the status flags if ``return_status``.

        ``pivoting="auto"`` re-integrates the systems the fused kernel flags (zero / small / non-finite pivot of
        the unpivoted LDL^T somewhere along the trajectory) from their initial state, stage by stage, with the
        dense pivoted-LU kernel (``forward_dynamics(..., pivoting="always")``), and marks them
        ``BIONC_STATUS_PIVOTED``; the default ``"never"`` returns the fused result with its flags and does not
        synchronise with the host.
        """
        if pivoting not in ("auto", "never"):
            raise ValueError("pivoting must be 'auto' or 'never'")
        q = _Batch(Q, self.nb_Q, self.device, "Q")
        qd = _Batch(Qdot, self.nb_Q, self.device, "Qdot")
        if q.B != qd.B:
            raise ValueError("Q and Qdot must have the same batch size")
        B = q.B
        if not (inplace and q.torch_in and qd.torch_in and q.t.data_ptr() == Q.data_ptr()):
            q.t, qd.t = q.t.clone(), qd.t.clone()
        start = (q.t.clone(), qd.t.clone()) if pivoting == "auto" else None
        h_steps = None
        hs = None
        if t is not None:
            tt = np.asarray(t, dtype=np.float64).reshape(-1)
            hs = tt[1:] - tt[:-1]
            n_steps = int(hs.size)
            h_steps = torch.from_numpy(np.ascontiguousarray(hs)).to(self.device)
            h = float(hs[0]) if hs.size else 0.0
        elif h is None or n_steps is None:
            raise ValueError("give either a time grid `t` or a constant step `h` with `n_steps`")
        traj = None
        if save_every is not None:
            traj = torch.empty((n_steps // save_every, B, 2 * self.nb_Q), dtype=torch.float64, device=self.device)
        lam = torch.empty((B, self.nb_holonomic_constraints), dtype=torch.float64, device=self.device) if return_lambdas else None
        status = torch.empty((B,), dtype=torch.int32, device=self.device)
        opt, keep = self._options(external_forces, stabilization, inertial_parameters, B, external_force_values,
                                  joint_forces=joint_generalized_forces if apply_joint_generalized_forces else None)
        if specialize is True or (specialize == "auto" and not self._no_jit and B * int(n_steps) >= 2_000_000 and self.nb_segments <= 4
                                  and (opt.fext or any(int(t) != 0 for t in self.table.blk_type))):
            try:
                with torch.cuda.device(self.device):
                    _lib.check(self._lib.bionc_cuda_model_specialize(self._h, C.byref(opt), 0, 1), "model_specialize")
            except _lib.BioncCudaError:
                if specialize is True:
                    raise
                self._no_jit = True  # no NVRTC / driver library here: the generic kernels serve every call
        with torch.cuda.device(self.device):
            if B > 0 and n_steps > 0:
                rc = self._lib.bionc_cuda_rk4(
                    self._h, q.t.data_ptr(), qd.t.data_ptr(), float(h), h_steps.data_ptr() if h_steps is not None else None,
                    int(n_steps), 1 if normalize else 0, C.byref(opt), traj.data_ptr() if traj is not None else None,
                    int(save_every or 1), lam.data_ptr() if lam is not None else None, status.data_ptr(), B, self._stream(),
                )
                _lib.check(rc, "rk4")
                if pivoting == "auto":
                    sel = torch.nonzero(status & 7).flatten()
                    if sel.numel() > 0:
                        steps = hs if hs is not None else np.full(int(n_steps), float(h))
                        self._rk4_stagewise_pivoted(sel, start, steps, normalize, external_forces, stabilization,
                                                    inertial_parameters, q.t, qd.t, traj, save_every, lam, status,
                                                    external_force_values)
            else:
                status.zero_()
        res = [q.out(q.t), q.out(qd.t)]
        if traj is not None:
            res.append(traj if q.torch_in else traj.cpu().numpy())
        if return_lambdas:
            res.append(q.out(lam, column=True))
        if return_status:
            res.append(q.out(status))
        return tuple(res)

    def _rk4_stagewise_pivoted(self, sel, start, steps, normalize, external_forces, stabilization, inertial_parameters,
                               Qout, Qdout, traj, save_every, lam, status, force_values=None):
        """Classic RK4 (utils/ode_solver.py:8-53) on the systems ``sel``, one dense pivoted-LU launch per stage."""
        yq, yv = start[0][sel].clone(), start[1][sel].clone()
        ip = None
        if inertial_parameters is not None:
            it = inertial_parameters if isinstance(inertial_parameters, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(inertial_parameters, dtype=np.float64))
            ip = it.to(self.device)[sel].contiguous()
        bits = torch.zeros((sel.numel(),), dtype=torch.int32, device=self.device)
        fv = None
        if force_values is not None:
            fv = force_values if isinstance(force_values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(force_values, dtype=np.float64))
            fv = fv.to(self.device)
            fv = fv[None] if fv.dim() == 3 else fv
            fv = fv[:, sel].contiguous()
        step_now = [0]

        def f(qq, vv):
            fvs = None if fv is None else fv[min(step_now[0], fv.shape[0] - 1)]
            a, l, s = self.forward_dynamics(qq, vv, external_forces=external_forces, stabilization=stabilization,
                                            inertial_parameters=ip, return_status=True, pivoting="always",
                                            external_force_values=fvs)
            bits.bitwise_or_(s & 3)
            return a, l

        n_steps = len(steps)
        for step in range(n_steps):
            h = float(steps[step])
            step_now[0] = step
            a1, lam1 = f(yq, yv)
            v2 = yv + 0.5 * h * a1
            a2, _ = f(yq + 0.5 * h * yv, v2)
            v3 = yv + 0.5 * h * a2
            a3, _ = f(yq + 0.5 * h * v2, v3)
            v4 = yv + h * a3
            a4, _ = f(yq + h * v3, v4)
            yq = yq + (h / 6.0) * (yv + 2.0 * v2 + 2.0 * v3 + v4)
            yv = yv + (h / 6.0) * (a1 + 2.0 * a2 + 2.0 * a3 + a4)
            if normalize:  # ode_solver.py:50-52 on model.normalized_coordinates
                y3 = yq.view(yq.shape[0], self.nb_segments, 4, 3)
                y3[:, :, 0] /= y3[:, :, 0].norm(dim=-1, keepdim=True)
                y3[:, :, 3] /= y3[:, :, 3].norm(dim=-1, keepdim=True)
            if traj is not None and (step + 1) % save_every == 0:
                traj[(step + 1) // save_every - 1, sel] = torch.cat([yq, yv], dim=1)
            if lam is not None and step == n_steps - 1:
                lam[sel] = lam1
        Qout[sel], Qdout[sel] = yq, yv
        bad = ~(torch.isfinite(yq).all(dim=1) & torch.isfinite(yv).all(dim=1))
        status[sel] = bits | 8 | (bad.to(torch.int32) * 2)

    # ------------------------------------------------------------------ host-buffer path (C ABI with host pointers)
    def rk4_host(self, Q: np.ndarray, Qdot: np.ndarray, h: float, n_steps: int, normalize: bool = True,
                 external_forces=None, stabilization: dict = None, status: np.ndarray | None = None,
                 external_force_values=None):
        """End-to-end call through ``bionc_cuda_rk4_host``: ``Q``/``Qdot`` are HOST arrays (B, n), float64,
        C-contiguous (pinned for full copy bandwidth), advanced in place; the library does H2D, the fused
        kernel and D2H and returns when the results are on the host."""
        for a in (Q, Qdot):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 2 and a.shape[1] == self.nb_Q):
                raise ValueError("rk4_host needs C-contiguous float64 host arrays of shape (B, 12N)")
        opt, keep = self._options(external_forces, stabilization, B=int(Q.shape[0]), force_values=external_force_values, host=True)
        _lib.check(
            self._lib.bionc_cuda_rk4_host(
                self._h, Q.ctypes.data, Qdot.ctypes.data, float(h), None, int(n_steps), 1 if normalize else 0, C.byref(opt),
                status.ctypes.data if status is not None else None, int(Q.shape[0]),
            ),
            "rk4_host",
        )
        return Q, Qdot

    def forward_dynamics_host(self, Q: np.ndarray, Qdot: np.ndarray, external_forces=None, stabilization: dict = None,
                              external_force_values=None):
        """End-to-end call through ``bionc_cuda_forward_dynamics_host`` with host arrays (B, n)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        Qdot = np.ascontiguousarray(Qdot, dtype=np.float64)
        B = Q.shape[0]
        qdd = np.empty_like(Q)
        lam = np.empty((B, self.nb_holonomic_constraints))
        status = np.empty((B,), dtype=np.int32)
        opt, keep = self._options(external_forces, stabilization, B=B, force_values=external_force_values, host=True)
        _lib.check(
            self._lib.bionc_cuda_forward_dynamics_host(
                self._h, Q.ctypes.data, Qdot.ctypes.data, C.byref(opt), qdd.ctypes.data, lam.ctypes.data, status.ctypes.data, B
            ),
            "forward_dynamics_host",
        )
        return qdd, lam, status

    def kernel_layout(self) -> dict:
        """How the model is laid out for the kernels (diagnostics): joint-level nodes, stored blocks, tree levels,
        systems per CTA and shared memory per CTA (the last two are known after the first launch)."""
        out = (C.c_int32 * 11)()
        _lib.check(self._lib.bionc_cuda_model_layout(self._h, out), "model_layout")
        keys = ("segments", "joint_rows", "nodes", "stored_blocks", "second_copies", "tree_levels", "row_storage_doubles",
                "zero_list", "systems_per_cta", "smem_bytes_per_cta", "serial_joint_level_schedule")
        return dict(zip(keys, [int(v) for v in out]))

    @staticmethod
    def launch_count() -> int:
        return int(_lib.load().bionc_cuda_launch_count())
