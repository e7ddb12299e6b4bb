"""``RK4`` and ``forward_integration`` with the reference's signatures (``bionc/utils/ode_solver.py``).

* :func:`RK4` is the generic, closure-driven integrator: it accepts any ``f(t, y, *args)`` and works
  on numpy arrays or torch tensors, for one state ``(d,)`` or a batch ``(B, d)``.  With a
  ``bionc_cuda`` model inside the closure every stage is one ``forward_dynamics`` kernel launch.
* :func:`forward_integration` uses the model's fused kernel (all four stages and all steps in one
  launch, :meth:`bionc_cuda.BiomechanicalModel.rk4`) -- a Python closure cannot be fused.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch


def RK4(t, f: Callable, y0, normalize_idx=None, args=()):
    """Classic fixed-step Runge-Kutta 4 on the grid ``t`` (ode_solver.py:8-53).

    Returns the states at every grid point: ``(d, T)`` for ``y0`` of shape ``(d,)`` (the reference's
    layout) and ``(B, d, T)`` for a batch ``(B, d)``.  ``normalize_idx`` is a sequence of index groups;
    after every step each group of every state is divided by its Euclidean norm (ode_solver.py:50-52).
    """
    is_torch = isinstance(y0, torch.Tensor)
    T = len(t)
    if is_torch:
        y = torch.zeros(tuple(y0.shape) + (T,), dtype=y0.dtype, device=y0.device)
        norm = lambda v: torch.linalg.vector_norm(v, dim=-1, keepdim=True)  # noqa: E731
    else:
        y0 = np.asarray(y0, dtype=np.float64)
        y = np.zeros(y0.shape + (T,))
        norm = lambda v: np.linalg.norm(v, axis=-1, keepdims=True)  # noqa: E731
    y[..., 0] = y0
    for i in range(T - 1):
        h = t[i + 1] - t[i]
        yi = y[..., i]
        k1 = f(t[i], yi, *args)
        k2 = f(t[i] + h / 2.0, yi + k1 * h / 2.0, *args)
        k3 = f(t[i] + h / 2.0, yi + k2 * h / 2.0, *args)
        k4 = f(t[i] + h, yi + k3 * h, *args)
        yn = yi + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        if normalize_idx is not None:
            for idx in normalize_idx:
                idx = list(idx)
                yn[..., idx] = yn[..., idx] / norm(yn[..., idx])
        y[..., i + 1] = yn
    return y


def forward_integration(model, Q_init, Qdot_init, t_final: float = 2, steps_per_second: int = 50):
    """ode_solver.py:56-122 on the CUDA backend: integrates ``[Q; Qdot]`` from 0 to ``t_final`` and
    returns ``(time_steps, all_states, dynamics)`` with ``all_states`` of shape ``(2n, T)`` for a
    single system (``(B, 2n, T)`` for a batch) and ``dynamics(t, states) -> (states_dot, lambdas)``.
    Raises ``ValueError`` when the initial state violates the rigid-body constraints like the reference."""
    phi = model.rigid_body_constraints(Q_init)
    phi_max = float(phi.max()) if not isinstance(phi, torch.Tensor) else float(phi.max().item())
    if phi_max > 1e-4:
        raise ValueError("The segment natural coordinates don't satisfy the rigid body constraint, at initial conditions.")
    time_steps = np.linspace(0, t_final, int(steps_per_second * t_final + 1))
    n = model.nb_Q

    def dynamics(t, states):
        q, qd = states[..., :n], states[..., n:]
        qddot, lambdas = model.forward_dynamics(q, qd)
        if isinstance(qddot, torch.Tensor):
            return torch.cat((qd, qddot.reshape(qd.shape)), dim=-1), lambdas
        return np.concatenate((np.asarray(qd).reshape(-1) if qddot.ndim == 2 and qddot.shape[-1] == 1 else qd,
                               qddot.reshape(-1) if qddot.ndim == 2 and qddot.shape[-1] == 1 else qddot), axis=-1), lambdas

    is_torch = isinstance(Q_init, torch.Tensor)
    Qf, Qdf, traj = model.rk4(Q_init, Qdot_init, t=time_steps, normalize=True, save_every=1)
    q0 = Q_init if is_torch else np.asarray(Q_init, dtype=np.float64)
    qd0 = Qdot_init if is_torch else np.asarray(Qdot_init, dtype=np.float64)
    single = q0.ndim == 1 or (q0.ndim == 2 and q0.shape[-1] == 1 and q0.shape[0] == n)
    if is_torch:
        y0 = torch.cat((q0.reshape(1, -1) if single else q0, qd0.reshape(1, -1) if single else qd0), dim=-1).to(traj.device)
        full = torch.cat((y0[None], traj), dim=0)  # (T, B, 2n)
        all_states = full.permute(1, 2, 0)
    else:
        y0 = np.concatenate((q0.reshape(1, -1) if single else q0, qd0.reshape(1, -1) if single else qd0), axis=-1)
        full = np.concatenate((y0[None], traj), axis=0)
        all_states = np.transpose(full, (1, 2, 0))
    if single:
        all_states = all_states[0]
    return time_steps, all_states, dynamics
