// Constrained forward dynamics of natural-coordinate models: one WARP per SEGMENT, one LANE per SYSTEM.
//
// System solved (reference bionc/bionc_numpy/biomechanical_model.py:404-426):
//     [G K^T; K 0] [Qdd; lambda] = [f; -bias]        G = blockdiag(M4_i (x) I3)
// The reference LU-factors the dense KKT matrix.  Here the same solution is obtained with a
// fill-reducing block elimination of the Schur complement S = K G^-1 K^T:
//   1. per segment (one thread per system, registers only): Sr_i = Kr_i G_i^-1 Kr_i^T (6x6) is factored LDL^T;
//      rigid-body rows of different segments never couple, so this is the block-diagonal leading
//      part of S under the ordering "all rigid rows first";
//   2. per constraint row touching the segment: q' = A_i k^T with the projected inverse mass
//      A_i = G_i^-1 - G_i^-1 Kr_i^T Sr_i^-1 Kr_i G_i^-1, and the segment's contribution
//      k_a . q'_b to the joint-level Schur complement S' (3x3 blocks in shared memory; closed form
//      gamma I - B M B^T for pairs of point-to-point blocks);
//   3. S' lambda_j = rhs' is factored as a sparse blocked LDL^T along a host-computed pattern
//      (minimum-degree ordering, fill, elimination-tree levels), block rows dealt to the warps of the CTA;
//   4. back-substitution per segment: F = f - Kj^T lambda_j, a = G^-1 F,
//      lambda_r = Sr^-1 (Kr a + bias_r), Qdd = a - G^-1 Kr^T lambda_r.
// LDL^T (no square roots) instead of Cholesky because the reference's mass matrix is not positive
// definite in general (SURVEY.md finding 1); a zero / non-finite pivot raises the status flag.
//
// Thread mapping: one CTA integrates 32 systems; WARP = SEGMENT, LANE = SYSTEM.  All lanes of a warp walk
// the same constraint-block lists (no divergence between segments), every shared-memory array is indexed
// [item][32 lanes] (conflict-free, immediate offsets) and the N warps of a CTA meet at block barriers
// (__syncthreads) between the phases; warps of lightly coupled segments simply wait there while the
// other resident CTAs use the issue slots.  Shared memory per CTA:
//   Qs, Qds  : N x 12 x 32   stage state (own segment at compile-time offsets, other segments at + 12*32*s)
//   S0, S1   : 9 nslot x 32, 9 nslot2 x 32   stored 3x3 blocks of the joint-level Schur matrix; a block that
//              two segments contribute to has a second copy written by the PARENT-role warp -> no atomics
//   r0, r1, lamv, dinv : 3 nb, 3 nb, 3 nb, 6 nb (x 32)   rhs (child / parent copies), y-hat -> lambda_j, pivots
//   knl      : N x rnl x 32   state-dependent vectors of the non-LIN3 rows (3 or 6 doubles per row and segment)
#pragma once
#include "model.cuh"

namespace bionc {

// A pivot (or 3x3 pivot determinant) this small relative to the diagonal scale raises the SMALL_PIVOT status
// bit: the unpivoted LDL^T may have lost accuracy (the reference's LU with partial pivoting would not).
constexpr double kSmallPivot = 1e-11;

// Reciprocal by the hardware seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps: 1 MUFU + 4 DFMA
// instead of the ~25-instruction IEEE division sequence; the result is within 1-2 ulp, far inside the
// 1e-10 parity tolerance.  x = 0 yields inf -> NaN, which the pivot checks flag before it is used.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// 1/sqrt(x) the same way (rsqrt.approx.ftz.f64 seed, two Newton steps)
__device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(0.5 * r, fma(-x * r, r, 1.0), r);
    r = fma(0.5 * r, fma(-x * r, r, 1.0), r);
    return r;
}

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// a + s * b
__device__ __forceinline__ V3 axpy(double s, V3 b, V3 a) { return V3{fma(s, b.x, a.x), fma(s, b.y, a.y), fma(s, b.z, a.z)}; }
__device__ __forceinline__ double comp(V3 a, int c) { return c == 0 ? a.x : (c == 1 ? a.y : a.z); }
__device__ __forceinline__ V3 unit(int c, double s) { return V3{c == 0 ? s : 0.0, c == 1 ? s : 0.0, c == 2 ? s : 0.0}; }

// A 12-vector as four 3-vectors on the slots (u, rp, rd, w).
struct Q12 {
    V3 s[4];
};

// index of W(t, s) in the packed upper triangle
__device__ __forceinline__ constexpr int widx(int t, int s) {
    return t <= s ? (t * 4 - t * (t - 1) / 2 + (s - t)) : (s * 4 - s * (s - 1) / 2 + (t - s));
}

// Per-lane view of one segment in one system.
struct SegCtx {
    double W[10];  // M4^-1
    double fg[4];
    V3 u, v, w;       // v = rp - rd
    double L[15];     // unit lower factor of Sr, packed rows: (1,0) (2,0) (2,1) (3,0) ...
    double dinv[6];
};

__device__ __forceinline__ constexpr int lidx(int r, int c) { return r * (r - 1) / 2 + c; }  // r > c

// x = G^-1 k  (W (x) I3 applied to a 12-vector)
__device__ __forceinline__ Q12 apply_W(const double (&W)[10], const Q12& k) {
    Q12 q;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        V3 acc = W[widx(t, 0)] * k.s[0];
#pragma unroll
        for (int s = 1; s < 4; ++s) acc = axpy(W[widx(t, s)], k.s[s], acc);
        q.s[t] = acc;
    }
    return q;
}

// Kr x for a 12-vector x (natural_segment.py:450-483 contracted): 6 numbers
__device__ __forceinline__ void kr_apply(const SegCtx& c, const Q12& x, double (&out)[6]) {
    const V3 xv = x.s[1] - x.s[2];
    out[0] = 2.0 * dot(c.u, x.s[0]);
    out[1] = dot(c.v, x.s[0]) + dot(c.u, xv);
    out[2] = dot(c.w, x.s[0]) + dot(c.u, x.s[3]);
    out[3] = 2.0 * dot(c.v, xv);
    out[4] = dot(c.w, xv) + dot(c.v, x.s[3]);
    out[5] = 2.0 * dot(c.w, x.s[3]);
}

// G^-1 Kr^T y for a 6-vector y
__device__ __forceinline__ Q12 ginv_krt(const SegCtx& c, const double (&y)[6]) {
    const V3 su = axpy(2.0 * y[0], c.u, axpy(y[1], c.v, y[2] * c.w));
    const V3 p = axpy(y[1], c.u, axpy(2.0 * y[3], c.v, y[4] * c.w));  // rp slot; rd slot is -p
    const V3 sw = axpy(y[2], c.u, axpy(y[4], c.v, (2.0 * y[5]) * c.w));
    Q12 x;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double wv = c.W[widx(t, 1)] - c.W[widx(t, 2)];
        x.s[t] = axpy(c.W[widx(t, 0)], su, axpy(wv, p, c.W[widx(t, 3)] * sw));
    }
    return x;
}

// solve Sr y = rhs in place with the LDL^T factors held in registers
__device__ __forceinline__ void sr_solve(const SegCtx& c, double (&y)[6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int r = k + 1; r < 6; ++r) y[r] = fma(-c.L[lidx(r, k)], y[k], y[r]);
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] *= c.dinv[k];
#pragma unroll
    for (int k = 5; k >= 0; --k)
#pragma unroll
        for (int r = k + 1; r < 6; ++r) y[k] = fma(-c.L[lidx(r, k)], y[r], y[k]);
}

// Build Sr = Kr G^-1 Kr^T from the Gram matrix of (u, v, w) and omega = E^T W E, E = [e_u, e_rp - e_rd, e_w],
// and factor it.  Row (a,b) of Kr is e_a (x) b + e_b (x) a, so
//   Sr[(a,b),(c,d)] = om_ac g_bd + om_ad g_bc + om_bc g_ad + om_bd g_ac.
__device__ __forceinline__ void sr_factor(SegCtx& c, int& status) {
    double g[3][3], om[3][3];
    g[0][0] = dot(c.u, c.u);
    g[0][1] = g[1][0] = dot(c.u, c.v);
    g[0][2] = g[2][0] = dot(c.u, c.w);
    g[1][1] = dot(c.v, c.v);
    g[1][2] = g[2][1] = dot(c.v, c.w);
    g[2][2] = dot(c.w, c.w);
    om[0][0] = c.W[widx(0, 0)];
    om[0][1] = om[1][0] = c.W[widx(0, 1)] - c.W[widx(0, 2)];
    om[0][2] = om[2][0] = c.W[widx(0, 3)];
    om[1][1] = c.W[widx(1, 1)] - 2.0 * c.W[widx(1, 2)] + c.W[widx(2, 2)];
    om[1][2] = om[2][1] = c.W[widx(1, 3)] - c.W[widx(2, 3)];
    om[2][2] = c.W[widx(3, 3)];
    constexpr int pa[6] = {0, 0, 0, 1, 1, 2};
    constexpr int pb[6] = {0, 1, 2, 1, 2, 2};
    double A[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) {
            const int a = pa[r], b = pb[r], cc = pa[q], d = pb[q];
            A[r][q] = fma(om[a][cc], g[b][d], fma(om[a][d], g[b][cc], fma(om[b][cc], g[a][d], om[b][d] * g[a][cc])));
        }
    double dmax = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) dmax = fmax(dmax, fabs(A[k][k]));
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double piv = A[k][k];
        if (!(fabs(piv) > 0.0) || !isfinite(piv)) status |= 1;
        else if (fabs(piv) < kSmallPivot * dmax) status |= 4;
        const double inv = fast_rcp(piv);
        c.dinv[k] = inv;
        double col[6];
#pragma unroll
        for (int r = k + 1; r < 6; ++r) {
            col[r] = A[r][k];
            c.L[lidx(r, k)] = col[r] * inv;
        }
#pragma unroll
        for (int r = k + 1; r < 6; ++r)
#pragma unroll
            for (int q = k + 1; q <= r; ++q) A[r][q] = fma(-c.L[lidx(r, k)], col[q], A[r][q]);
    }
}

// ---------------------------------------------------------------------------------------------
// warp-level context.  All shared-memory pointers are pre-offset for this lane, so that the hot
// code addresses them with compile-time multiples of a stride:
//   Qo/Qdo     own segment's stage state, slot element k at Qo[k * 32]          ([N][12][32 lanes])
//   Qsys/Qdsys segment 0 of this lane's system; segment s, element k at Qsys[(12 s + k) * 32]
//   S0/S1      joint-level Schur matrix as 3x3 blocks: the block (I, J), I >= J, lives in storage slot
//              p = slot_tab[I nb + J] (structural non-zeros + fill only), element e = 3 r + c at
//              S0[(9 p + e) * spw]; S1 holds the second exclusive-writer copy of the blocks that can
//              receive contributions from two segments (slot2_tab)           ([9 nslot][spw])
//   r0/r1      right-hand side / y-hat, row j at r0[j * spw]                    ([3 nb][spw])
//   dinv       inverse diagonal blocks, 6 unique entries of block K at dinv[(6 K + i) * spw]
//   knl        state-dependent vectors of the non-LIN3 rows (d1; d2 too for sphere-on-plane parent sides),
//              double offset q element i at knl[(q + i) * 32]
// ---------------------------------------------------------------------------------------------
struct WarpCtx {
    const DevHeader* hdr;
    const DevSeg* seg;
    const DevBlk* blk;
    const DevSide* side;
    const DevFext* fext;
    double *Qo, *Qdo, *Qsys, *Qdsys;
    double *S0, *S1, *r0, *r1, *lamv, *dinv, *knl;
    // symbolic factorisation tables (host, capi.cu symbolic_analysis)
    const int *slot_tab, *slot2_tab, *zero_list, *comb_list, *unit_list, *colptr, *colrow, *lvlptr, *lvlpiv, *owner, *pivmask, *pivwriter, *nscale;
    int N, mj, nb;
    int segi;  // this warp's segment
    int lane;  // this lane's system inside the CTA
};


// linear form over a segment's published state: sum_t a[t] * X[segment s, slot t]
template <int S>
__device__ __forceinline__ V3 lin_form(const WarpCtx& w, const double* Xsys, int s, const double* a) {
    V3 r = mk(0.0, 0.0, 0.0);
    const double* X = Xsys + 12 * S * s;
#pragma unroll
    for (int t = 0; t < 4; ++t) r = axpy(a[t], mk(X[(3 * t) * S], X[(3 * t + 1) * S], X[(3 * t + 2) * S]), r);
    return r;
}

struct DynOpts {
    int use_stab;
    double alpha, beta;
};

// Jacobian row (restricted to this lane's segment) of a non-LIN3 block.  Every such row has the form
//     k = a1 (x) d1  [+ a2 (x) d2]
// with constant 4-coefficient vectors a (block parameters) and state-dependent 3-vectors d.  nonlin_row
// evaluates d1, d2 (these 6 doubles are what is kept in shared memory per row and lane) and -- on the
// child side -- the row's right-hand-side term bias (+ alpha Phi + beta K Qdot when stabilised);
// nonlin_k rebuilds the 12-vector from them.
// DOT : joints.py:203-208, :222-235, :252-282 ; joints_with_ground.py:155-162, :169-178
// CLEN: joints.py:1300-1382        SOP: joints.py:657-748 (Jacobian blocks as coded: swapped)
__device__ __forceinline__ Q12 nonlin_k(const DevBlk& b, int role, V3 d1, V3 d2) {
    const double* p = b.p;
    Q12 k;
    if (b.type == DOT) {
        const double* a = role == CHILD ? p + 4 : p;
#pragma unroll
        for (int t = 0; t < 4; ++t) k.s[t] = a[t] * d1;
    } else if (b.type == CLEN) {
        const double* a = role == CHILD ? p + 4 : p;
        const double sg = role == CHILD ? -2.0 : 2.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) k.s[t] = (sg * a[t]) * d1;
    } else {  // SOP
        if (role == CHILD) {
#pragma unroll
            for (int t = 0; t < 4; ++t) k.s[t] = p[t] * d1;  // n^T N_P, stored in the child's columns
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) k.s[t] = axpy(-p[4 + t], d1, p[8 + t] * d2);  // stored in the parent's columns
        }
    }
    return k;
}

template <int S>
__device__ __forceinline__ void nonlin_row(const WarpCtx& w, const DevBlk& b, int role, const DynOpts& o, V3& d1, V3& d2,
                                           double& bterm) {
    const double* p = b.p;
    bterm = 0.0;
    d2 = mk(0.0, 0.0, 0.0);
    if (b.type == DOT) {
        const V3 Dc = lin_form<S>(w, w.Qsys, b.child, p + 4);
        V3 Dp, Dpd = mk(0.0, 0.0, 0.0);
        if (b.parent >= 0) {
            Dp = lin_form<S>(w, w.Qsys, b.parent, p);
        } else {
            Dp = mk(p[0], p[1], p[2]);
        }
        if (role == CHILD) {
            d1 = Dp;
            const V3 Dcd = lin_form<S>(w, w.Qdsys, b.child, p + 4);
            if (b.parent >= 0) {
                Dpd = lin_form<S>(w, w.Qdsys, b.parent, p);
                bterm = 2.0 * dot(Dpd, Dcd);
            }
            if (o.use_stab) bterm += o.alpha * (dot(Dp, Dc) - p[8]) + o.beta * (dot(Dpd, Dc) + dot(Dp, Dcd));
        } else {
            d1 = Dc;
        }
    } else if (b.type == CLEN) {
        const V3 d = lin_form<S>(w, w.Qsys, b.parent, p) - lin_form<S>(w, w.Qsys, b.child, p + 4);
        d1 = d;
        if (role == CHILD) {
            const V3 dd = lin_form<S>(w, w.Qdsys, b.parent, p) - lin_form<S>(w, w.Qdsys, b.child, p + 4);
            bterm = 2.0 * dot(dd, dd);
            if (o.use_stab) bterm += o.alpha * (dot(d, d) - p[8]) + o.beta * (2.0 * dot(d, dd));
        }
    } else {  // SOP
        const V3 n = lin_form<S>(w, w.Qsys, b.child, p + 8);
        d1 = n;
        if (role == CHILD) {
            const V3 P = lin_form<S>(w, w.Qsys, b.parent, p), A = lin_form<S>(w, w.Qsys, b.child, p + 4);
            const V3 Pd = lin_form<S>(w, w.Qdsys, b.parent, p), Ad = lin_form<S>(w, w.Qdsys, b.child, p + 4);
            const V3 nd = lin_form<S>(w, w.Qdsys, b.child, p + 8);
            bterm = 2.0 * dot(Pd, nd) - 2.0 * dot(nd, Ad);
            if (o.use_stab) {
                // K Qdot with the blocks where the reference stores them:
                // parent cols: -n^T N_A + (P-A)^T N_n applied to Qdot_parent; child cols: n^T N_P applied to Qdot_child
                const V3 Apd = lin_form<S>(w, w.Qdsys, b.parent, p + 4), npd = lin_form<S>(w, w.Qdsys, b.parent, p + 8);
                const V3 Pcd = lin_form<S>(w, w.Qdsys, b.child, p);
                const double kqd = -dot(n, Apd) + dot(P - A, npd) + dot(n, Pcd);
                bterm += o.alpha * (dot(P - A, n) - p[12]) + o.beta * kqd;
            }
        } else {
            d2 = lin_form<S>(w, w.Qsys, b.parent, p) - lin_form<S>(w, w.Qsys, b.child, p + 4);
        }
    }
}

// the stored state-dependent vectors of a non-LIN3 row of this lane: d1 at offset q (in doubles), d2 behind
// it for the only two-term row kind (sphere-on-plane, parent side)
__device__ __forceinline__ bool nonlin_two_terms(const DevBlk& b, int role) { return b.type == SOP && role == PARENT; }
template <int S>
__device__ __forceinline__ Q12 load_nonlin_k(const WarpCtx& w, const DevBlk& b, int role, int q) {
    const double* src = w.knl + q * S;
    const V3 d1 = mk(src[0], src[S], src[2 * S]);
    const V3 d2 = nonlin_two_terms(b, role) ? mk(src[3 * S], src[4 * S], src[5 * S]) : mk(0.0, 0.0, 0.0);
    return nonlin_k(b, role, d1, d2);
}

// External forces on this lane's segment (external_force.py:143-180 and the four force classes).
__device__ __forceinline__ Q12 external_forces(const WarpCtx& w, const DevSeg& sg, const SegCtx& c, V3 rp) {
    Q12 f;
#pragma unroll
    for (int t = 0; t < 4; ++t) f.s[t] = mk(0.0, 0.0, 0.0);
    for (int e = sg.fext_begin; e < sg.fext_end; ++e) {
        const DevFext& fe = w.fext[e];
        V3 tau = mk(fe.p[0], fe.p[1], fe.p[2]), F = mk(fe.p[3], fe.p[4], fe.p[5]);
        const V3 pt = mk(fe.p[6], fe.p[7], fe.p[8]);
        if (fe.kind == 3) {  // in local: rotate by [u v w] Tinv (external_force_in_local.py:85-91)
            const double* T = fe.p + 9;
            const V3 c0 = axpy(T[0], c.u, axpy(T[3], c.v, T[6] * c.w));
            const V3 c1 = axpy(T[1], c.u, axpy(T[4], c.v, T[7] * c.w));
            const V3 c2 = axpy(T[2], c.u, axpy(T[5], c.v, T[8] * c.w));
            const V3 Fg = axpy(F.x, c0, axpy(F.y, c1, F.z * c2));
            const V3 tg = axpy(tau.x, c0, axpy(tau.y, c1, tau.z * c2));
            F = Fg;
            tau = tg;
        }
        if (fe.kind == 1 || fe.kind == 3) {  // local application point (external_force_global_local_point.py:68-100)
            // rp - P = -(pt0 u + pt1 (rp - rd) + pt2 w)
            const V3 lever = mk(0.0, 0.0, 0.0) - axpy(pt.x, c.u, axpy(pt.y, c.v, pt.z * c.w));
            tau = tau + cross(lever, F);
        } else if (fe.kind == 2) {  // global application point (external_force_global.py:67-96)
            tau = tau + cross(rp - pt, F);
        }
        // pseudo interpolation matrix (natural_coordinates.py:132-158): solve [w x u, u x v, -v x w] x = tau
        const V3 b0 = cross(c.w, c.u), b1 = cross(c.u, c.v), b2 = cross(c.w, c.v);  // -v x w = w x v
        const double det = dot(b0, cross(b1, b2));
        const double idet = fast_rcp(det);
        const double x0 = dot(tau, cross(b1, b2)) * idet;
        const double x1 = dot(b0, cross(tau, b2)) * idet;
        const double x2 = dot(b0, cross(b1, tau)) * idet;
        f.s[0] = axpy(x1, c.v, f.s[0]);
        f.s[1] = f.s[1] + F - x2 * c.w;
        f.s[2] = axpy(x2, c.w, f.s[2]);
        f.s[3] = axpy(x0, c.u, f.s[3]);
    }
    return f;
}

// this lane's own segment, slot t, from / to the published stage state (Xo = w.Qo or w.Qdo)
template <int S>
__device__ __forceinline__ V3 own3(const double* Xo, int t) {
    return mk(Xo[(3 * t) * S], Xo[(3 * t + 1) * S], Xo[(3 * t + 2) * S]);
}
template <int S>
__device__ __forceinline__ void publish3(double* Xo, int t, V3 v) {
    Xo[(3 * t) * S] = v.x, Xo[(3 * t + 1) * S] = v.y, Xo[(3 * t + 2) * S] = v.z;
}

// rigid-body right-hand side term: bias_r (+ alpha Phi_r + beta Kr Qdot)   (natural_segment.py:429-448, :508-545)
template <int S>
__device__ __forceinline__ void rigid_bterm(const WarpCtx& w, const SegCtx& c, const DevSeg& sg, const DynOpts& o,
                                            double (&b)[6]) {
    Q12 Qd;
#pragma unroll
    for (int t = 0; t < 4; ++t) Qd.s[t] = own3<S>(w.Qdo, t);
    const V3 ud = Qd.s[0], vd = Qd.s[1] - Qd.s[2], wd = Qd.s[3];
    b[0] = 2.0 * dot(ud, ud);
    b[1] = 2.0 * dot(ud, vd);
    b[2] = 2.0 * dot(ud, wd);
    b[3] = 2.0 * dot(vd, vd);
    b[4] = 2.0 * dot(vd, wd);
    b[5] = 2.0 * dot(wd, wd);
    if (o.use_stab) {
        double kqd[6];
        kr_apply(c, Qd, kqd);
        const double phi[6] = {dot(c.u, c.u) - 1.0,         dot(c.u, c.v) - sg.phic[0], dot(c.u, c.w) - sg.phic[1],
                               dot(c.v, c.v) - sg.phic[2], dot(c.v, c.w) - sg.phic[3], dot(c.w, c.w) - 1.0};
#pragma unroll
        for (int i = 0; i < 6; ++i) b[i] += o.alpha * phi[i] + o.beta * kqd[i];
    }
}

// three right-hand sides at once (the three rows of a LIN3 block): 3-way instruction-level parallelism
__device__ __forceinline__ void sr_solve3(const SegCtx& c, double (&y)[6][3]) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int r = k + 1; r < 6; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) y[r][j] = fma(-c.L[lidx(r, k)], y[k][j], y[r][j]);
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) y[k][j] *= c.dinv[k];
#pragma unroll
    for (int k = 5; k >= 0; --k)
#pragma unroll
        for (int r = k + 1; r < 6; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) y[k][j] = fma(-c.L[lidx(r, k)], y[r][j], y[k][j]);
}

// signed coefficient vector of a LIN3 block restricted to one side: rows are  as (x) e_c
__device__ __forceinline__ void lin3_coefs(const DevBlk& b, int role, double (&as)[4]) {
    const double* a = b.p + (role == CHILD ? 4 : 0);
    const int flip = role == CHILD ? (int)0x80000000 : 0;  // sign flip on the integer pipe: the FP64 pipe is the scarce one
#pragma unroll
    for (int t = 0; t < 4; ++t) as[t] = __hiloint2double(__double2hiint(a[t]) ^ flip, __double2loint(a[t]));
}
// beta = W as, and its (u, v, w) reduction b3 = E^T beta with E = [e_u, e_rp - e_rd, e_w]
__device__ __forceinline__ void lin3_beta(const SegCtx& c, const double (&as)[4], double (&beta)[4], double (&b3)[3]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) acc = fma(c.W[widx(t, s)], as[s], acc);
        beta[t] = acc;
    }
    b3[0] = beta[0], b3[1] = beta[1] - beta[2], b3[2] = beta[3];
}

// ---------------------------------------------------------------------------------------------
// One forward-dynamics evaluation for the whole warp.
// Precondition: every lane has published its segment's stage state (Q, Qdot) in w.Qs / w.Qds.
// Every lane receives Qddot of its segment; lambda (holonomic order) goes to lam_out if non-null.
// Every thread of the CTA must call this (it contains __syncthreads).
//
// GENERAL=false is the lean instantiation for models whose joints are all point-to-point (LIN3:
// spherical / ground spherical / weld halves), without external forces or stabilisation -- the
// lower-limb and full-body trees.  GENERAL=true handles everything.
//
// For two LIN3 block sides a, b on one segment (rows as_a (x) e_c', as_b (x) e_c) the contribution to
// the joint-level Schur complement has the closed form
//     S'_ab = (as_a . W as_b) I3 - B M_ab B^T,   M_ab = Gamma(a3)^T Sr^-1 Gamma(b3),   B = [u v w]
// with Gamma(b3) the 6x3 matrix [[2bu,0,0],[bv,bu,0],[bw,0,bu],[0,2bv,0],[0,bw,bv],[0,0,2bw]]
// (Kr G^-1 (as_b (x) e_c)^T = Gamma(b3) B[c,:]^T), so the three rows of a block share one
// three-right-hand-side solve Z_b = Sr^-1 Gamma(b3).
// ---------------------------------------------------------------------------------------------
// 3x3 helpers for the blocked joint-level factorisation (blocks are stored row-major, element e = 3 r + c)
template <int spw>
__device__ __forceinline__ void load9(const double* p, double (&A)[9]) {
#pragma unroll
    for (int e = 0; e < 9; ++e) A[e] = p[e * spw];
}
// Effective block (I, J) of the joint-level matrix: storage slot plus, where two segments contribute, the
// second exclusive-writer copy.  The copies are never merged in memory: every reader adds them, every
// update goes to the first copy.
template <int spw>
__device__ __forceinline__ void load_block(const WarpCtx& w, int idx, double (&A)[9]) {
    load9<spw>(w.S0 + 9 * w.slot_tab[idx] * spw, A);
    const int s2 = w.slot2_tab[idx];
    if (s2 >= 0) {
        const double* p2 = w.S1 + 9 * s2 * spw;
#pragma unroll
        for (int e = 0; e < 9; ++e) A[e] += p2[e * spw];
    }
}

// inverse of a symmetric 3x3 given by its lower triangle (adjugate); di = [i00 i10 i11 i20 i21 i22]
__device__ __forceinline__ void sym3_inverse(double a00, double a10, double a11, double a20, double a21, double a22,
                                             double (&di)[6], int& status) {
    const double c00 = a11 * a22 - a21 * a21;
    const double c10 = a20 * a21 - a10 * a22;
    const double c20 = a10 * a21 - a20 * a11;
    const double det = fma(a00, c00, fma(a10, c10, a20 * c20));
    const double dmax = fmax(fabs(a00), fmax(fabs(a11), fabs(a22)));
    if (!(fabs(det) > 0.0) || !isfinite(det)) status |= 1;
    else if (fabs(det) < kSmallPivot * dmax * dmax * dmax) status |= 4;
    const double id = fast_rcp(det);
    di[0] = c00 * id;
    di[1] = c10 * id;
    di[2] = (a00 * a22 - a20 * a20) * id;
    di[3] = c20 * id;
    di[4] = (a10 * a20 - a00 * a21) * id;
    di[5] = (a00 * a11 - a10 * a10) * id;
}
// y = Dinv x for the symmetric 3x3 Dinv stored as 6 entries
__device__ __forceinline__ V3 sym3_apply(const double (&di)[6], V3 x) {
    return mk(fma(di[0], x.x, fma(di[1], x.y, di[3] * x.z)), fma(di[1], x.x, fma(di[2], x.y, di[4] * x.z)),
              fma(di[3], x.x, fma(di[4], x.y, di[5] * x.z)));
}

// Pivot K of the blocked LDL^T: invert D_K, y-hat_K = D_K^-1 r_K, then for the rows I of column K that this warp
// owns (ALL: every row) the fused forward substitution and the Schur updates S_IJ -= S_IK D_K^-1 S_JK^T.
template <bool ALL, int spw>
__device__ __forceinline__ void eliminate_pivot(const WarpCtx& w, int K, int& status) {
    const int nb = w.nb;
    const int c0 = w.colptr[K], c1 = w.colptr[K + 1];
    double D[9], di[6];
    load_block<spw>(w, K * nb + K, D);
    sym3_inverse(D[0], D[3], D[4], D[6], D[7], D[8], di, status);
    // a pivot whose whole diagonal has cancelled against the node's nominal scale is a redundant node (e.g. the same
    // joint declared twice): its remainder is rounding noise that the determinant test, relative to the block's own
    // scale, cannot see
    if (fmax(fabs(D[0]), fmax(fabs(D[4]), fabs(D[8]))) < kSmallPivot * (double)__int_as_float(w.nscale[K])) status |= 4;
    const V3 yh = sym3_apply(di, mk(w.r0[(3 * K) * spw] + w.r1[(3 * K) * spw], w.r0[(3 * K + 1) * spw] + w.r1[(3 * K + 1) * spw],
                                    w.r0[(3 * K + 2) * spw] + w.r1[(3 * K + 2) * spw]));
    if (ALL || w.segi == w.pivwriter[K]) {
#pragma unroll
        for (int i = 0; i < 6; ++i) w.dinv[(6 * K + i) * spw] = di[i];
        w.lamv[(3 * K) * spw] = yh.x, w.lamv[(3 * K + 1) * spw] = yh.y, w.lamv[(3 * K + 2) * spw] = yh.z;  // y-hat = D^-1 y
    }
#pragma unroll 1
    for (int ci = c0; ci < c1; ++ci) {
        const int I = w.colrow[ci];
        if (!ALL && w.owner[I] != w.segi) continue;
        double A[9], T[9];
        load_block<spw>(w, I * nb + K, A);
        double* ri = w.r0 + 3 * I * spw;  // fused forward substitution
        ri[0] -= fma(A[0], yh.x, fma(A[1], yh.y, A[2] * yh.z));
        ri[spw] -= fma(A[3], yh.x, fma(A[4], yh.y, A[5] * yh.z));
        ri[2 * spw] -= fma(A[6], yh.x, fma(A[7], yh.y, A[8] * yh.z));
#pragma unroll
        for (int r = 0; r < 3; ++r) {  // T = A Dinv
            const V3 t = sym3_apply(di, mk(A[3 * r], A[3 * r + 1], A[3 * r + 2]));
            T[3 * r] = t.x, T[3 * r + 1] = t.y, T[3 * r + 2] = t.z;
        }
#pragma unroll 1
        for (int cj = c0; cj <= ci; ++cj) {  // every J <= I of the same column: (I, J) is in the pattern
            const int J = w.colrow[cj];
            double Bj[9];
            load_block<spw>(w, J * nb + K, Bj);
            double* Sij = w.S0 + 9 * w.slot_tab[I * nb + J] * spw;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double upd = fma(T[3 * r], Bj[3 * q], fma(T[3 * r + 1], Bj[3 * q + 1], T[3 * r + 2] * Bj[3 * q + 2]));
                    Sij[(3 * r + q) * spw] -= upd;
                }
        }
    }
}

// lambda_K = y-hat_K - D_K^-1 sum_{I in column K} S_IK^T lambda_I   (in place in lamv)
template <int spw>
__device__ __forceinline__ void back_substitute(const WarpCtx& w, int K) {
    const int c0 = w.colptr[K], c1 = w.colptr[K + 1];
    if (c0 == c1) return;  // root: lambda = y-hat
    V3 acc = mk(0.0, 0.0, 0.0);
#pragma unroll 1
    for (int ci = c0; ci < c1; ++ci) {
        const int I = w.colrow[ci];
        double A[9];
        load_block<spw>(w, I * w.nb + K, A);
        const V3 li = mk(w.lamv[(3 * I) * spw], w.lamv[(3 * I + 1) * spw], w.lamv[(3 * I + 2) * spw]);
        acc = axpy(li.x, mk(A[0], A[1], A[2]), axpy(li.y, mk(A[3], A[4], A[5]), axpy(li.z, mk(A[6], A[7], A[8]), acc)));
    }
    double di[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) di[i] = w.dinv[(6 * K + i) * spw];
    const V3 corr = sym3_apply(di, acc);
    w.lamv[(3 * K) * spw] -= corr.x;
    w.lamv[(3 * K + 1) * spw] -= corr.y;
    w.lamv[(3 * K + 2) * spw] -= corr.z;
}

// Joint-level system of three nodes shaped like an arrow: nodes 0 and 1 are leaves coupled only with the root 2
// (the lower limb: hip and ankle against the knee).  The two leaves are eliminated concurrently by the first two
// warps; their updates of the root go to the two exclusive-writer copies of D_2 and r_2, so no write is shared.
// After one barrier both warps invert the root (cheaper than a second barrier) and each finishes its own leaf:
// lambda_leaf = y-hat_leaf - T^T lambda_2 with T = S_2,leaf D_leaf^-1 kept in the block's first copy.
template <int spw>
__device__ __forceinline__ void solve_two_leaves(const WarpCtx& w, int& status) {
    constexpr int nb = 3;
    if (w.segi < 2) {
        const int K = w.segi;
        double D[9], di[6], A[9], T[9];
        load_block<spw>(w, K * nb + K, D);
        sym3_inverse(D[0], D[3], D[4], D[6], D[7], D[8], di, status);
        if (fmax(fabs(D[0]), fmax(fabs(D[4]), fabs(D[8]))) < kSmallPivot * (double)__int_as_float(w.nscale[K])) status |= 4;
        const V3 yh = sym3_apply(di, mk(w.r0[(3 * K) * spw] + w.r1[(3 * K) * spw], w.r0[(3 * K + 1) * spw] + w.r1[(3 * K + 1) * spw],
                                        w.r0[(3 * K + 2) * spw] + w.r1[(3 * K + 2) * spw]));
        w.lamv[(3 * K) * spw] = yh.x, w.lamv[(3 * K + 1) * spw] = yh.y, w.lamv[(3 * K + 2) * spw] = yh.z;
        load_block<spw>(w, 2 * nb + K, A);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const V3 t = sym3_apply(di, mk(A[3 * r], A[3 * r + 1], A[3 * r + 2]));
            T[3 * r] = t.x, T[3 * r + 1] = t.y, T[3 * r + 2] = t.z;
        }
        double* Sd = K == 0 ? w.S0 + 9 * w.slot_tab[2 * nb + 2] * spw : w.S1 + 9 * w.slot2_tab[2 * nb + 2] * spw;
        double* rr = (K == 0 ? w.r0 : w.r1) + 6 * spw;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                Sd[(3 * r + q) * spw] -= fma(T[3 * r], A[3 * q], fma(T[3 * r + 1], A[3 * q + 1], T[3 * r + 2] * A[3 * q + 2]));
            rr[r * spw] -= fma(A[3 * r], yh.x, fma(A[3 * r + 1], yh.y, A[3 * r + 2] * yh.z));
        }
        double* Sa = w.S0 + 9 * w.slot_tab[2 * nb + K] * spw;
#pragma unroll
        for (int e = 0; e < 9; ++e) Sa[e * spw] = T[e];
    }
    __syncthreads();
    if (w.segi < 2) {
        const int K = w.segi;
        double D[9], di[6], T[9];
        load_block<spw>(w, 2 * nb + 2, D);
        sym3_inverse(D[0], D[3], D[4], D[6], D[7], D[8], di, status);
        if (fmax(fabs(D[0]), fmax(fabs(D[4]), fabs(D[8]))) < kSmallPivot * (double)__int_as_float(w.nscale[2])) status |= 4;
        const V3 l2 = sym3_apply(di, mk(w.r0[6 * spw] + w.r1[6 * spw], w.r0[7 * spw] + w.r1[7 * spw], w.r0[8 * spw] + w.r1[8 * spw]));
        if (K == 0) w.lamv[6 * spw] = l2.x, w.lamv[7 * spw] = l2.y, w.lamv[8 * spw] = l2.z;
        load9<spw>(w.S0 + 9 * w.slot_tab[2 * nb + K] * spw, T);
        w.lamv[(3 * K) * spw] -= fma(T[0], l2.x, fma(T[3], l2.y, T[6] * l2.z));
        w.lamv[(3 * K + 1) * spw] -= fma(T[1], l2.x, fma(T[4], l2.y, T[7] * l2.z));
        w.lamv[(3 * K + 2) * spw] -= fma(T[2], l2.x, fma(T[5], l2.y, T[8] * l2.z));
    }
}

template <bool GENERAL, int S>
__device__ __forceinline__ void eval_forward_dynamics(const WarpCtx& w, SegCtx& c, const DynOpts& o, Q12& Qdd,
                                                      double* lam_out, int& status) {
    constexpr int spw = S;
    const DevSeg& sg = w.seg[w.segi];

    // ---- B. segment-local factorisation
    const V3 rp = own3<S>(w.Qo, 1);
    c.u = own3<S>(w.Qo, 0), c.v = rp - own3<S>(w.Qo, 2), c.w = own3<S>(w.Qo, 3);
    sr_factor(c, status);
    Q12 f;  // natural forces: gravity + external
    if (GENERAL && sg.fext_end > sg.fext_begin) {
        f = external_forces(w, sg, c, rp);
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) f.s[t] = mk(0.0, 0.0, 0.0);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) f.s[t].z += c.fg[t];
    DynOpts oo = o;
    if (!GENERAL) oo.use_stab = 0;

    // Two passes over the same segment solve  x = A_i f + (particular part):
    //   pass 0 with f = applied forces gives the free acceleration abar that feeds the joint-level system,
    //   pass 1 with f - Kj^T lambda_j gives Qddot and lambda_r.
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        Q12 x;
        double z[6];
        {
            const Q12 a = apply_W(c.W, f);
            kr_apply(c, a, z);
            double br[6];  // recomputed in both passes rather than kept alive across the joint-level phases
            rigid_bterm<S>(w, c, sg, oo, br);
#pragma unroll
            for (int i = 0; i < 6; ++i) z[i] += br[i];
            sr_solve(c, z);
            const Q12 corr = ginv_krt(c, z);
#pragma unroll
            for (int t = 0; t < 4; ++t) x.s[t] = a.s[t] - corr.s[t];
        }
        if (pass == 1 || w.mj == 0) {
            Qdd = x;
            if (lam_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 6; ++i) lam_out[sg.rigid_row + i] = z[i];  // lambda_r
            }
            break;
        }
        const int nb = w.nb, mjp = 3 * nb;
        const DevHeader* hd = w.hdr;

        // ---- C. zero what the assembly does not overwrite (pure fill blocks, blocks of nodes made of single
        //         rows); padding rows get a unit diagonal and a zero right-hand side.  Lean models have
        //         neither, so this phase and its barrier vanish for them.
        if (hd->nzero > 0 || hd->nunit > 0) {
            for (int zi = w.segi; zi < hd->nzero; zi += w.N) {
                const int z = w.zero_list[zi];
                double* dst = z >= 0 ? w.S0 + 9 * z * spw : w.S1 + 9 * (-z - 1) * spw;
#pragma unroll
                for (int e = 0; e < 9; ++e) dst[e * spw] = 0.0;
            }
            __syncthreads();
            if (w.segi == 0) {
                for (int it = 0; it < hd->nunit; ++it) w.S0[w.unit_list[it] * spw] = 1.0;
                for (int r = w.mj; r < mjp; ++r) w.r0[r * spw] = 0.0, w.r1[r * spw] = 0.0;
            }
        }

        // ---- D. right-hand side rhs' = Kj abar + b_j (copy `role` of each row is written by exactly one
        //         lane) and the explicit rows of non-LIN3 blocks
        {
            for (int e = sg.side_begin; e < sg.side_end; ++e) {
                const DevSide sd = w.side[e];
                const DevBlk& b = w.blk[sd.blk];
                double* rr = (sd.role == CHILD ? w.r0 : w.r1) + b.jrow * spw;
                if (!GENERAL || b.type == LIN3) {
                    double as[4];
                    lin3_coefs(b, sd.role, as);
                    V3 r = mk(0.0, 0.0, 0.0);
#pragma unroll
                    for (int t = 0; t < 4; ++t) r = axpy(as[t], x.s[t], r);
                    if (GENERAL && sd.role == CHILD && o.use_stab) {
                        // Phi = c0 + N_p Q_p - N_c Q_c ; K Qdot = N_p Qdot_p - N_c Qdot_c
                        V3 phi = mk(b.p[8], b.p[9], b.p[10]) - lin_form<S>(w, w.Qsys, b.child, b.p + 4);
                        V3 kqd = mk(0.0, 0.0, 0.0) - lin_form<S>(w, w.Qdsys, b.child, b.p + 4);
                        if (b.parent >= 0) {
                            phi = phi + lin_form<S>(w, w.Qsys, b.parent, b.p);
                            kqd = kqd + lin_form<S>(w, w.Qdsys, b.parent, b.p);
                        }
                        r = axpy(o.alpha, phi, axpy(o.beta, kqd, r));
                    }
                    rr[0] = r.x, rr[spw] = r.y, rr[2 * spw] = r.z;
                    if (b.parent < 0) {  // no parent lane: the child also settles the second copy
                        double* r1p = w.r1 + b.jrow * spw;
                        r1p[0] = 0.0, r1p[spw] = 0.0, r1p[2 * spw] = 0.0;
                    }
                } else {
                    V3 d1, d2;
                    double bterm;
                    nonlin_row<S>(w, b, sd.role, o, d1, d2, bterm);
                    const Q12 k = nonlin_k(b, sd.role, d1, d2);
                    double acc = bterm;
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc += dot(k.s[t], x.s[t]);
                    double* dst = w.knl + sd.nl_slot * S;
                    dst[0] = d1.x, dst[S] = d1.y, dst[2 * S] = d1.z;
                    if (nonlin_two_terms(b, sd.role)) dst[3 * S] = d2.x, dst[4 * S] = d2.y, dst[5 * S] = d2.z;
                    rr[0] = acc;
                    if (b.parent < 0) w.r1[b.jrow * spw] = 0.0;
                }
            }
        }

        // ---- E1. LIN3 x LIN3 pairs on this segment: S'_ab = gamma I3 - B M_ab B^T, one 3x3 block each
        {
            for (int e = sg.side_begin; e < sg.side_end; ++e) {
                const DevSide sd = w.side[e];
                const DevBlk& b = w.blk[sd.blk];
                if (GENERAL && b.type != LIN3) continue;
                double asb[4], betab[4], b3[3];
                lin3_coefs(b, sd.role, asb);
                lin3_beta(c, asb, betab, b3);
                double Z[6][3];  // Sr^-1 Gamma(b3)
                Z[0][0] = 2.0 * b3[0], Z[0][1] = 0.0, Z[0][2] = 0.0;
                Z[1][0] = b3[1], Z[1][1] = b3[0], Z[1][2] = 0.0;
                Z[2][0] = b3[2], Z[2][1] = 0.0, Z[2][2] = b3[0];
                Z[3][0] = 0.0, Z[3][1] = 2.0 * b3[1], Z[3][2] = 0.0;
                Z[4][0] = 0.0, Z[4][1] = b3[2], Z[4][2] = b3[1];
                Z[5][0] = 0.0, Z[5][1] = 0.0, Z[5][2] = 2.0 * b3[2];
                sr_solve3(c, Z);
                const int Jb = b.jrow / 3;
                for (int e2 = e; e2 < sg.side_end; ++e2) {
                    const DevSide sa = w.side[e2];
                    const DevBlk& a = w.blk[sa.blk];
                    if (GENERAL && a.type != LIN3) continue;
                    double asa[4], betaa[4], a3[3];
                    lin3_coefs(a, sa.role, asa);
                    lin3_beta(c, asa, betaa, a3);
                    double gamma = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) gamma = fma(asa[t], betab[t], gamma);
                    // M = Gamma(a3)^T Z (3x3), P = B M (columns), block = gamma I - P B^T
                    V3 Pc[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double m0 = fma(2.0 * a3[0], Z[0][j], fma(a3[1], Z[1][j], a3[2] * Z[2][j]));
                        const double m1 = fma(a3[0], Z[1][j], fma(2.0 * a3[1], Z[3][j], a3[2] * Z[4][j]));
                        const double m2 = fma(a3[0], Z[2][j], fma(a3[1], Z[4][j], (2.0 * a3[2]) * Z[5][j]));
                        Pc[j] = axpy(m0, c.u, axpy(m1, c.v, m2 * c.w));
                    }
                    // element (row r of a, row cc of b) = gamma delta - sum_j Pc[j][r] * B[cc][j],  B[cc] = (u_cc, v_cc, w_cc)
                    const V3 col0 = mk(0, 0, 0) - axpy(c.u.x, Pc[0], axpy(c.v.x, Pc[1], c.w.x * Pc[2]));
                    const V3 col1 = mk(0, 0, 0) - axpy(c.u.y, Pc[0], axpy(c.v.y, Pc[1], c.w.y * Pc[2]));
                    const V3 col2 = mk(0, 0, 0) - axpy(c.u.z, Pc[0], axpy(c.v.z, Pc[1], c.w.z * Pc[2]));
                    const int Ia = a.jrow / 3;
                    if (Ia >= Jb) {  // block (Ia, Jb): rows of a, columns of b
                        const int s2 = w.slot2_tab[Ia * nb + Jb];
                        double* Sc = (sa.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Ia * nb + Jb] * spw;
                        Sc[0 * spw] = col0.x + gamma, Sc[1 * spw] = col1.x, Sc[2 * spw] = col2.x;
                        Sc[3 * spw] = col0.y, Sc[4 * spw] = col1.y + gamma, Sc[5 * spw] = col2.y;
                        Sc[6 * spw] = col0.z, Sc[7 * spw] = col1.z, Sc[8 * spw] = col2.z + gamma;
                    } else {  // block (Jb, Ia): transposed; the copy follows this lane's role in the larger node b
                        const int s2 = w.slot2_tab[Jb * nb + Ia];
                        double* Sc = (sd.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Jb * nb + Ia] * spw;
                        Sc[0 * spw] = col0.x + gamma, Sc[1 * spw] = col0.y, Sc[2 * spw] = col0.z;
                        Sc[3 * spw] = col1.x, Sc[4 * spw] = col1.y + gamma, Sc[5 * spw] = col1.z;
                        Sc[6 * spw] = col2.x, Sc[7 * spw] = col2.y, Sc[8 * spw] = col2.z + gamma;
                    }
                }
            }
        }

        // ---- E2. pairs with at least one non-LIN3 row: explicit q' = A_i k_b^T of the non-LIN3 row b
        if (GENERAL) {
            for (int e = sg.side_begin; e < sg.side_end; ++e) {
                const DevSide sd = w.side[e];
                const DevBlk& b = w.blk[sd.blk];
                if (b.type == LIN3) continue;
                const Q12 kb = load_nonlin_k<S>(w, b, sd.role, sd.nl_slot);
                Q12 q;
                {
                    double y[6];
                    q = apply_W(c.W, kb);
                    kr_apply(c, q, y);
                    sr_solve(c, y);
                    const Q12 corr = ginv_krt(c, y);
#pragma unroll
                    for (int t = 0; t < 4; ++t) q.s[t] = q.s[t] - corr.s[t];
                }
                const int Ib = b.jrow / 3, rb = b.jrow - 3 * Ib;
                for (int e2 = sg.side_begin; e2 < sg.side_end; ++e2) {
                    const DevSide sa = w.side[e2];
                    const DevBlk& a = w.blk[sa.blk];
                    // the copy is chosen by this lane's role in the block holding the larger row
                    if (a.type == LIN3) {
                        double asa[4];
                        lin3_coefs(a, sa.role, asa);
                        V3 r = mk(0.0, 0.0, 0.0);
#pragma unroll
                        for (int t = 0; t < 4; ++t) r = axpy(asa[t], q.s[t], r);
                        const int Ia = a.jrow / 3;  // LIN3 blocks are whole nodes, never Ib
                        if (Ia > Ib) {  // rows of a, column rb
                            const int s2 = w.slot2_tab[Ia * nb + Ib];
                            double* Sc = ((sa.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Ia * nb + Ib] * spw) + rb * spw;
                            Sc[0] = r.x, Sc[3 * spw] = r.y, Sc[6 * spw] = r.z;
                        } else {  // row rb, columns of a
                            const int s2 = w.slot2_tab[Ib * nb + Ia];
                            double* Sc = ((sd.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Ib * nb + Ia] * spw) + 3 * rb * spw;
                            Sc[0] = r.x, Sc[spw] = r.y, Sc[2 * spw] = r.z;
                        }
                    } else if (e2 >= e) {  // every pair of single rows once
                        const Q12 ka = load_nonlin_k<S>(w, a, sa.role, sa.nl_slot);
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < 4; ++t) acc += dot(ka.s[t], q.s[t]);
                        const int Ia = a.jrow / 3, ra = a.jrow - 3 * Ia;
                        // the larger ROW decides the block orientation and whose role picks the copy
                        const bool a_larger = a.jrow >= b.jrow;
                        const int Ihi = a_larger ? Ia : Ib, Ilo = a_larger ? Ib : Ia;
                        const int rhi = a_larger ? ra : rb, rlo = a_larger ? rb : ra;
                        const int role = a_larger ? sa.role : sd.role;
                        const int s2 = w.slot2_tab[Ihi * nb + Ilo];
                        double* Sc = (role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Ihi * nb + Ilo] * spw;
                        Sc[(3 * rhi + rlo) * spw] = acc;
                        if (Ihi == Ilo && rhi != rlo) Sc[(3 * rlo + rhi) * spw] = acc;  // keep diagonal blocks symmetric
                    }
                }
            }
        }

        // ---- F. sparse blocked LDL^T of S' in place with 3x3 pivots along the host-computed pattern.  The forward
        //         substitution is fused in (r_I -= S_IK D_K^-1 r_K); S_IK is kept unscaled (L_IK = S_IK D_K^-1 acts
        //         through y-hat).
        //    G. backward substitution: lambda_K = y-hat_K - D_K^-1 sum_I S_IK^T lambda_I.
        if (hd->serial_solve == 2) {
            __syncthreads();  // closes the assembly
            solve_two_leaves<S>(w, status);
        } else if (hd->serial_solve != 0) {
            // serial schedule (host decision: the elimination offers no parallelism across warps, e.g. the lower
            // limb's hip / ankle pivots both update the knee block): the first warp factors and solves everything
            // between two barriers -- no barrier per level, no pivot inverted twice.
            __syncthreads();
            if (w.segi == 0) {
#pragma unroll 1
                for (int K = 0; K < nb; ++K) eliminate_pivot<true, S>(w, K, status);
#pragma unroll 1
                for (int K = nb - 1; K >= 0; --K) back_substitute<S>(w, K);
            }
        } else {
            // level by level of the elimination tree (pivots of one level are independent: one barrier per level, the
            // first one also closes the assembly).  A block row I is always handled by warp owner[I], so concurrent
            // pivots never write the same block.
#pragma unroll 1
            for (int lv = 0; lv < hd->nlevel; ++lv) {
                __syncthreads();
#pragma unroll 1
                for (int pi = w.lvlptr[lv]; pi < w.lvlptr[lv + 1]; ++pi) {
                    const int K = w.lvlpiv[pi];
                    if (((w.pivmask[K] >> w.segi) & 1) == 0) continue;  // neither the pivot's writer nor the owner of a row of column K
                    eliminate_pivot<false, S>(w, K, status);
                }
            }
            // Back substitution level by level from the roots down, the pivots of a level spread over the warps (they
            // only read lambda of higher levels); the barrier of a level also closes the y-hat / inverse pivots that the
            // factorisation stored above it.
#pragma unroll 1
            for (int lv = hd->nlevel - 2; lv >= 0; --lv) {  // the last level holds roots only: lambda = y-hat already
                __syncthreads();
#pragma unroll 1
                for (int pi = w.lvlptr[lv] + w.segi; pi < w.lvlptr[lv + 1]; pi += w.N) back_substitute<S>(w, w.lvlpiv[pi]);
            }
        }
        __syncthreads();

        // ---- H. f <- f - Kj^T lambda_j for the second pass.  Without external forces f is just gravity: it is
        //         rebuilt here instead of being kept alive (24 registers) across the joint-level phases.
        if (!GENERAL) {
#pragma unroll
            for (int t = 0; t < 4; ++t) f.s[t] = mk(0.0, 0.0, c.fg[t]);
        }
        {
            for (int e = sg.side_begin; e < sg.side_end; ++e) {
                const DevSide sd = w.side[e];
                const DevBlk& b = w.blk[sd.blk];
                const double* lj = w.lamv + b.jrow * spw;
                if (!GENERAL || b.type == LIN3) {
                    double as[4];
                    lin3_coefs(b, sd.role, as);
                    const V3 lam = mk(lj[0], lj[spw], lj[2 * spw]);
#pragma unroll
                    for (int t = 0; t < 4; ++t) f.s[t] = axpy(-as[t], lam, f.s[t]);
                    if (lam_out != nullptr && sd.role == CHILD) {
                        lam_out[b.row_h + 0] = lam.x, lam_out[b.row_h + 1] = lam.y, lam_out[b.row_h + 2] = lam.z;
                    }
                } else {
                    const double lam = lj[0];
                    const Q12 k = load_nonlin_k<S>(w, b, sd.role, sd.nl_slot);
#pragma unroll
                    for (int t = 0; t < 4; ++t) f.s[t] = axpy(-lam, k.s[t], f.s[t]);
                    if (lam_out != nullptr && sd.role == CHILD) lam_out[b.row_h] = lam;
                }
            }
        }
    }
    // No barrier here: lambda_j lives in its own array, so the next evaluation may start assembling while
    // slower warps still read it; the barriers inside the next evaluation order everything else.
}

}  // namespace bionc
