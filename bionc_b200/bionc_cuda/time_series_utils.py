"""Constraint residuals over a time series, the functions of ``bionc_numpy/time_series_utils.py`` on the CUDA
backend.  The reference walks the frames of ``Q`` (n, frames) in a Python loop; here every frame of every system
is one row of a batch and the norms are reduced on the device (``BIONC_EVAL_RESIDUALS``), so the trajectory
written by ``model.rk4(..., save_every=k)`` -- (frames, B, 2n) on the device -- can be checked in place.

Layouts: ``Q`` may be (n, frames) like the reference (``frames_last=True``, the default for 2-D numpy input whose
first axis is n), or (..., n) with any leading axes (frames, systems).  A trajectory (frames, B, 2n) is accepted as
is: the velocities are sliced off."""

from __future__ import annotations

import numpy as np


def _rows(model, Q, frames_last):
    import torch

    n = model.nb_Q
    is_torch = isinstance(Q, torch.Tensor)
    x = Q if is_torch else np.asarray(Q, dtype=np.float64)
    if frames_last is None:
        frames_last = (not is_torch) and x.ndim == 2 and x.shape[0] == n and x.shape[1] != n
    if frames_last:
        x = x.T if not is_torch else x.transpose(0, 1)
    if x.shape[-1] == 2 * n:
        x = x[..., :n]
    if x.shape[-1] != n:
        raise ValueError(f"last axis must have length {n} (Q) or {2 * n} ([Q, Qdot]); got {tuple(x.shape)}")
    lead = tuple(x.shape[:-1])
    flat = x.reshape(-1, n)
    return flat, lead


def total_rigid_body_constraints(model, Q, frames_last=None):
    """time_series_utils.py:24-26: ``sqrt(sum(Phi_r^2))`` per frame -> shape of Q without its coordinate axis."""
    flat, lead = _rows(model, Q, frames_last)
    return model.constraint_residual_norms(flat)[:, 0].reshape(lead)


def total_joint_constraints(model, Q, frames_last=None):
    """time_series_utils.py:34-36"""
    flat, lead = _rows(model, Q, frames_last)
    return model.constraint_residual_norms(flat)[:, 1].reshape(lead)


def rigid_body_constraints(model, Q, frames_last=None):
    """time_series_utils.py:29-31: the residual vectors themselves, (..., 6N) -- (6N, frames) for reference-shaped input."""
    flat, lead = _rows(model, Q, frames_last)
    out = model.rigid_body_constraints(flat).reshape(lead + (-1,))
    return out.T if _reference_shaped(model, Q, frames_last) else out


def joint_constraints(model, Q, frames_last=None):
    """time_series_utils.py:39-41"""
    flat, lead = _rows(model, Q, frames_last)
    out = model.joint_constraints(flat).reshape(lead + (-1,))
    return out.T if _reference_shaped(model, Q, frames_last) else out


def _reference_shaped(model, Q, frames_last):
    if frames_last is not None:
        return bool(frames_last)
    return isinstance(Q, np.ndarray) and Q.ndim == 2 and Q.shape[0] == model.nb_Q and Q.shape[1] != model.nb_Q


class TimeSeriesUtils:
    """Namespace mirror of ``bionc_numpy.time_series_utils.TimeSeriesUtils``."""

    total_rigid_body_constraints = staticmethod(total_rigid_body_constraints)
    rigid_body_constraints = staticmethod(rigid_body_constraints)
    total_joint_constraints = staticmethod(total_joint_constraints)
    joint_constraints = staticmethod(joint_constraints)
