"""Algorithmic FLOP / byte counts of the hot path, computed from a compiled
model table.  These are the numerators of ``bench.py``'s roofline block: they describe the
*reference-shaped* Schur-complement algorithm with a dense factorisation of the m x m Schur
complement, independent of how the kernels actually organise the work (the kernels exploit the
block structure and execute fewer FLOPs than this count).

    F_eval = F_asm + F_T + F_S + F_fac + F_solve + F_rhs + F_back         (1 FMA = 2 FLOPs)
      F_asm   = 40 N + 2 nnz(K)
      F_T     = 96 * (number of non-zero 12-wide row blocks of K)           M4^-1 (x) I3 per block
      F_S     = 24 * sum_{i<=j} (number of segments rows i and j share)     S = K G^-1 K^T
      F_fac   = m^3/3 + m^2/2
      F_solve = 2 m^2
      F_rhs   = F_back = 96 N + 2 nnz(K)
    F_step = 4 F_eval + 16 n           one RK4 step;  bytes/step = 2 * 2n * 8 (read + write Q, Qdot)
"""

from __future__ import annotations

import numpy as np

from ..bionc_cuda.model_table import BLK_CLEN, BLK_DOT, BLK_LIN3, BLK_SOP, ModelTable


def _row_supports(table: ModelTable):
    """For every holonomic row: {segment: structural nnz in that segment's 12 columns}."""
    rows = [dict() for _ in range(table.nb_holonomic_constraints)]
    rigid_nnz = (3, 9, 6, 6, 9, 3)  # natural_segment.py:465-481
    for i in range(table.nb_segments):
        for k in range(6):
            rows[int(table.seg_rigid_row[i]) + k][i] = rigid_nnz[k]
    for b in range(table.nb_blocks):
        t, p, c, r0 = int(table.blk_type[b]), int(table.blk_parent[b]), int(table.blk_child[b]), int(table.blk_row_h[b])
        P = table.blk_param[b]
        nz = lambda a: int(np.count_nonzero(a))  # noqa: E731
        if t == BLK_LIN3:
            for k in range(3):
                rows[r0 + k][c] = nz(P[4:8])
                if p >= 0:
                    rows[r0 + k][p] = nz(P[0:4])
        elif t == BLK_DOT:
            if p >= 0:
                rows[r0][c] = 3 * nz(P[4:8])
                rows[r0][p] = 3 * nz(P[0:4])
            else:  # constant ground axis: zeros of the axis are structural zeros of the row
                rows[r0][c] = nz(P[0:3]) * nz(P[4:8])
        elif t == BLK_CLEN:
            rows[r0][c] = 3 * nz(P[4:8])
            rows[r0][p] = 3 * nz(P[0:4])
        elif t == BLK_SOP:
            rows[r0][p] = 3 * nz(np.abs(P[4:8]) + np.abs(P[8:12]))  # as coded: -n N_A + (P-A) N_n in the parent columns
            rows[r0][c] = 3 * nz(P[0:4])
    return rows


def flops_per_eval(table: ModelTable) -> float:
    rows = _row_supports(table)
    N, m = table.nb_segments, table.nb_holonomic_constraints
    nnz = sum(sum(r.values()) for r in rows)
    blocks = sum(len(r) for r in rows)
    shared = 0
    for i in range(m):
        si = set(rows[i])
        for j in range(i, m):
            shared += len(si & set(rows[j]))
    f_asm = 40 * N + 2 * nnz
    f_t = 96 * blocks
    f_s = 24 * shared
    f_fac = m**3 / 3 + m**2 / 2
    f_solve = 2 * m**2
    f_rhs = f_back = 96 * N + 2 * nnz
    return float(f_asm + f_t + f_s + f_fac + f_solve + f_rhs + f_back)


def flops_per_step(table: ModelTable) -> float:
    return 4.0 * flops_per_eval(table) + 16.0 * table.nb_Q


def bytes_per_step(table: ModelTable) -> float:
    """Algorithmic HBM traffic of one RK4 system-step: read + write of Q and Qdot."""
    return 2.0 * 2.0 * table.nb_Q * 8.0
