// Constrained forward dynamics of natural-coordinate models: one WARP per SEGMENT, one LANE per SYSTEM.
//
// System solved (reference bionc/bionc_numpy/biomechanical_model.py:404-426):
//     [G K^T; K 0] [Qdd; lambda] = [f; -bias]        G = blockdiag(M4_i (x) I3)
// The reference LU-factors the dense KKT matrix.  The kernels obtain the same solution by eliminating the
// rigid-body rows of every segment in closed form (null-space form) and the joint rows with a sparse block LDL^T:
//
//   1. Rigid-body rows, per segment (one thread per system, registers only).  The six rows of Kr are the gradients of
//      u.u, u.v, u.w, v.v, v.w, w.w (natural_segment.py:450-483); for ANY invertible basis B = [u v w] (drifted
//      or not) the null space of Kr is the set of rigid motions  dQ = (al x u, t, t - al x v, al x w), and
//      [A_u A_v A_w] = -1/2 B^-T Bsym solves  Kr q_p = -b_r  when Bsym is the symmetric matrix of the six
//      right-hand sides b_r (bias: Bsym = 2 Ud^T Ud).  So  Qdd_i = T_i (al_i, tc_i) + q_p,i  with six unknowns per
//      segment: the angular acceleration al and the acceleration tc of the centre of mass.  Projecting
//      G_i Qdd_i + Kr^T lambda_r + Kj^T lambda_j = f_i  with T_i^T removes lambda_r and leaves Newton-Euler:
//          m tc = F - sum_rows fdir lambda          Ic al = tau_c - eta - sum_rows n lambda
//      m = total mass, Ic = tr(Y) I - Y with Y = B N3 B^T and N3 = J - m c c^T the second-moment matrix about
//      the centre of mass in natural coordinates, eta = sum_ij N3_ij B_i x A_j, (tau_c, F) the applied wrench about
//      the centre of mass (gravity: pure force).  Only 3x3 matrices are inverted; G^-1 never appears, so an
//      indefinite M4 costs nothing as long as m != 0 and Ic is regular.
//   2. Every joint row restricted to a segment, k = a1 (x) d1 [+ a2 (x) d2] (a: constant 4 coefficients over
//      (u, rp, rd, w), d: 3-vector), acts on (al, tc) through the 6-vector (n, fdir) = (rho' x d, sigma d) with
//      sigma = a_rp + a_rd, rho' = B (a~ - sigma c), a~ = (a_u, -a_rd, a_w), and on q_p through kappa = d . A (a~ - sigma c).
//      The segment's contribution to the joint-level Schur complement S' is
//          S'_ab = fdir_a . fdir_b / m + n_a . Ic^-1 n_b
//      (3x3 blocks gamma I - [rho_a]x Ic^-1 [rho_b]x for two point-to-point sides), its right-hand side
//          n . al_free + fdir . tc_free + kappa (+ bias + Baumgarte terms, added by the child side).
//   3. S' lambda_j = rhs' is factored as a sparse blocked LDL^T along a host-computed pattern
//      (minimum-degree ordering, fill, elimination-tree levels), block rows dealt to the warps of the CTA.
//   4. Back-substitution per segment: (al, tc) from Newton-Euler with the joint multipliers, Qdd = T (al, tc) + q_p;
//      lambda_r, when asked for, from  Kr^T lambda_r = f - Kj^T lambda_j - G Qdd  through B^-1.
// The result is algebraically identical to the reference's KKT solution (tests: 1e-10 on Qdd and lambda); a zero /
// non-finite determinant raises the status flag.
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
//   knl      : N x rnl x 32   the 6-vector (n, f_dir) of every non-LIN3 row touching the segment (6 doubles per row)
#pragma once
#include "model.cuh"
#include "tmem.cuh"

// Model-specialised builds (BIONC_SPEC, spec.cuh): the model tables are constants of the translation unit and a warp's
// segment is a template parameter, so the loops over the tables have compile-time trip counts.  BIONC_TAB_LOOP marks
// them: unrolled completely there (every table entry folds into the instruction stream), left to the compiler otherwise;
// BIONC_TAB_LOOP1 marks the ones the generic build keeps rolled.
#if defined(BIONC_SPEC) && defined(BIONC_SPEC_ROLL_SOLVE)
// ... except the factorisation of the joint-level system when the generated unit asks to keep its loops: unrolled, the
// elimination of a 7-node system is 150 KB of straight-line code per CTA against a 32 KB instruction cache
#define BIONC_TAB_LOOP _Pragma("unroll")
#define BIONC_TAB_LOOP1 _Pragma("unroll 1")
#elif defined(BIONC_SPEC)
#define BIONC_TAB_LOOP _Pragma("unroll")
#define BIONC_TAB_LOOP1 _Pragma("unroll")
#else
#define BIONC_TAB_LOOP
#define BIONC_TAB_LOOP1 _Pragma("unroll 1")
#endif

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
// acc + a x b and acc - a x b in six fused multiply-adds (a standalone cross product plus an addition costs nine FP64
// issue slots: every DMUL / DADD takes the same slot as a DFMA)
__device__ __forceinline__ V3 cross_acc(V3 a, V3 b, V3 acc) {
    return V3{fma(a.y, b.z, fma(-a.z, b.y, acc.x)), fma(a.z, b.x, fma(-a.x, b.z, acc.y)), fma(a.x, b.y, fma(-a.y, b.x, acc.z))};
}
__device__ __forceinline__ V3 cross_sub(V3 acc, V3 a, V3 b) { return cross_acc(b, a, acc); }
// a + s * b
__device__ __forceinline__ V3 axpy(double s, V3 b, V3 a) { return V3{fma(s, b.x, a.x), fma(s, b.y, a.y), fma(s, b.z, a.z)}; }
__device__ __forceinline__ double comp(V3 a, int c) { return c == 0 ? a.x : (c == 1 ? a.y : a.z); }
__device__ __forceinline__ V3 unit(int c, double s) { return V3{c == 0 ? s : 0.0, c == 1 ? s : 0.0, c == 2 ? s : 0.0}; }

// A 12-vector as four 3-vectors on the slots (u, rp, rd, w).
struct Q12 {
    V3 s[4];
};

// index of entry (t, s) of a symmetric 4x4 in the packed upper triangle (uu ur ud uw rr rd rw dd dw ww)
__device__ __forceinline__ constexpr int widx(int t, int s) {
    return t <= s ? (t * 4 - t * (t - 1) / 2 + (s - t)) : (s * 4 - s * (s - 1) / 2 + (t - s));
}

// Per-lane view of one segment in one system.
struct SegCtx {
    // Constants of the segment, 11 doubles read from shared memory where they are used (no registers held across the
    // evaluation): kc[i * ks] = N3 (6: J - m c c^T packed xx xy xz yy yz zz, the second-moment matrix about the centre
    // of mass in natural coordinates), natural centre of mass (3), m, 1 / m.  Model inertia: the segment's table entry
    // (ks = 1, one broadcast address per warp); per-system inertial parameters (config 5): this lane's column of the
    // CTA's constant area (ks = systems per CTA), filled by load_constants.
    const double* kc;
    int ks;
    __device__ __forceinline__ double N3(int i) const { return kc[i * ks]; }
    __device__ __forceinline__ double cn(int i) const { return kc[(6 + i) * ks]; }
    __device__ __forceinline__ double m() const { return kc[9 * ks]; }
    __device__ __forceinline__ double minv() const { return kc[10 * ks]; }
    // state-dependent, set by segment_setup
    V3 u, v, w;        // v = rp - rd
    V3 rp;             // proximal point (read again only by the multiplier output of the system-row layout)
    V3 Au, Av, Aw;     // particular accelerations of u, v, w:  [Au Av Aw] = B^-T S,  S = -1/2 Bsym
    double Ii[6];      // inverse of the inertia tensor about the centre of mass (sym3_inverse layout: 00 10 11 20 21 22)
    V3 tau0, F0;       // applied wrench about the centre of mass, gyroscopic term eta already subtracted from tau0
};

// sum_i x[i] * (u, v, w)[i]
__device__ __forceinline__ V3 in_basis(const SegCtx& c, double x0, double x1, double x2) {
    return axpy(x0, c.u, axpy(x1, c.v, x2 * c.w));
}
// same combination of the particular accelerations
__device__ __forceinline__ V3 in_accel(const SegCtx& c, double x0, double x1, double x2) {
    return axpy(x0, c.Au, axpy(x1, c.Av, x2 * c.Aw));
}
// acc + sum_i x[i] * (Au, Av, Aw)[i]
__device__ __forceinline__ V3 in_accel_acc(const SegCtx& c, double x0, double x1, double x2, V3 acc) {
    return axpy(x0, c.Au, axpy(x1, c.Av, axpy(x2, c.Aw, acc)));
}

// A constant coefficient vector a over the slots (u, rp, rd, w) as seen from the centre of mass:
// sum_s a[s] Qdd_s = sigma tc + al x rho' + kappa',  sigma = a_rp + a_rd,  rho' = B r,  kappa' = A r,
// r = (a_u, -a_rd, a_w) - sigma c.
struct Lever {
    double sigma;
    double r[3];
};
__device__ __forceinline__ Lever lever_of(const SegCtx& c, const double (&a)[4]) {
    Lever l;
    l.sigma = a[1] + a[2];
    l.r[0] = fma(-l.sigma, c.cn(0), a[0]);
    l.r[1] = fma(-l.sigma, c.cn(1), -a[2]);
    l.r[2] = fma(-l.sigma, c.cn(2), a[3]);
    return l;
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
    const DevFextTab* ftab;  // this call's external forces (shared-memory copy), nullptr without forces
    const DevFext* fext;
    const double* fvals;     // per-system (torque, force) values of this lane's system at the current step, or nullptr
    double *Qo, *Qdo, *Qsys, *Qdsys;
    double *S0, *S1, *r0, *r1, *lamv, *dinv, *knl;
    double* kcs;  // per-system constants of this lane's segment, element i at kcs[i * spw] (allocated only for such calls)
    // symbolic factorisation tables (host, capi.cu symbolic_analysis)
    const int *slot_rt, *slot_tab, *slot2_tab, *zero_list, *comb_list, *unit_list, *colptr, *colrow, *lvlptr, *lvlpiv, *owner, *pivmask, *pivwriter, *nscale;
    int N, mj, nb;
    int segi;  // this warp's segment
    int lane;  // this lane's system inside the CTA
    bool persys;  // per-system inertial parameters: the levers of the model's centre of mass are corrected per system
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
    int poison;  // debug (bionc_cuda_debug_poison): bit mask of working-set regions filled with NaN before the first evaluation
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
// the 12-vector of a non-LIN3 row restricted to this lane's segment, rebuilt from the published state (multiplier
// output path only: the hot path keeps the row's 6-vector instead, load_row6)
template <int S>
__device__ __forceinline__ Q12 rebuild_nonlin_k(const WarpCtx& w, const DevBlk& b, int role) {
    V3 d1, d2;
    double bterm;
    DynOpts none;
    none.use_stab = 0, none.poison = 0, none.alpha = 0.0, none.beta = 0.0;
    nonlin_row<S>(w, b, role, none, d1, d2, bterm);
    return nonlin_k(b, role, d1, d2);
}

// External forces (external_force.py:143-180 and the four force classes).  The force table lives outside the model
// blob: DevFextTab | DevFext[nf], sorted by segment (begin[s] .. begin[s+1]).  `vals` (optional) holds per-system
// values: 6 doubles (torque, force) per force, indexed by the force's position in the caller's list (DevFext::slot).
// fext_transport returns the force and its torque about rp.  Kind 4 is the reaction of a joint generalized force
// (generalized_force.py:98-129: the force acts on the joint's child at its proximal point, `rp_other`; the parent gets
// the opposite wrench transported to its own proximal point, external_force_global_on_proximal.py:117-147 as coded).
__device__ __forceinline__ void fext_transport(const DevFext& fe, const double* __restrict__ vals, V3 u, V3 v, V3 w, V3 rp,
                                               V3 rp_other, V3& tau, V3& F) {
    if (vals != nullptr) {
        const double* x = vals + 6 * fe.slot;
        tau = mk(x[0], x[1], x[2]), F = mk(x[3], x[4], x[5]);
    } else {
        tau = mk(fe.p[0], fe.p[1], fe.p[2]), F = mk(fe.p[3], fe.p[4], fe.p[5]);
    }
    const V3 pt = mk(fe.p[6], fe.p[7], fe.p[8]);
    if (fe.kind == 3) {  // in local: rotate by [u v w] Tinv (external_force_in_local.py:85-91)
        const double* T = fe.p + 9;
        const V3 c0 = axpy(T[0], u, axpy(T[3], v, T[6] * w));
        const V3 c1 = axpy(T[1], u, axpy(T[4], v, T[7] * w));
        const V3 c2 = axpy(T[2], u, axpy(T[5], v, T[8] * w));
        const V3 Fg = axpy(F.x, c0, axpy(F.y, c1, F.z * c2));
        const V3 tg = axpy(tau.x, c0, axpy(tau.y, c1, tau.z * c2));
        F = Fg;
        tau = tg;
    }
    if (fe.kind == 1 || fe.kind == 3) {  // local application point (external_force_global_local_point.py:68-100)
        // rp - P = -(pt0 u + pt1 (rp - rd) + pt2 w)
        const V3 lever = mk(0.0, 0.0, 0.0) - axpy(pt.x, u, axpy(pt.y, v, pt.z * w));
        tau = tau + cross(lever, F);
    } else if (fe.kind == 2) {  // global application point (external_force_global.py:67-96)
        tau = tau + cross(rp - pt, F);
    } else if (fe.kind == 4) {  // -(tau + (rp - rp_other) x F, F)
        tau = cross_acc(rp - rp_other, F, tau);
        tau = mk(-tau.x, -tau.y, -tau.z), F = mk(-F.x, -F.y, -F.z);
    }
}

// Natural external forces of one segment as the reference forms them: the torque goes through the pseudo
// interpolation matrix (natural_coordinates.py:132-158), f = N_prox^T F + N*^T tau.  Used by the FEXT evaluator, the
// dense kernel and the lambda_r output; the dynamics only need the wrench (below).
template <typename RpOf>  // RpOf(segment) -> proximal point of another segment of the same system (kind 4 only)
__device__ __forceinline__ Q12 external_forces(const DevFext* __restrict__ fext, int begin, int end,
                                               const double* __restrict__ vals, V3 u, V3 v, V3 w, V3 rp, RpOf rp_of) {
    Q12 f;
#pragma unroll
    for (int t = 0; t < 4; ++t) f.s[t] = mk(0.0, 0.0, 0.0);
    for (int e = begin; e < end; ++e) {
        V3 tau, F;
        fext_transport(fext[e], vals, u, v, w, rp, fext[e].kind == 4 ? rp_of(fext[e].other) : rp, tau, F);
        // solve [w x u, u x v, -v x w] x = tau
        const V3 b0 = cross(w, u), b1 = cross(u, v), b2 = cross(w, v);  // -v x w = w x v
        const double det = dot(b0, cross(b1, b2));
        const double idet = fast_rcp(det);
        const double x0 = dot(tau, cross(b1, b2)) * idet;
        const double x1 = dot(b0, cross(tau, b2)) * idet;
        const double x2 = dot(b0, cross(b1, tau)) * idet;
        f.s[0] = axpy(x1, v, f.s[0]);
        f.s[1] = f.s[1] + F - x2 * w;
        f.s[2] = axpy(x2, w, f.s[2]);
        f.s[3] = axpy(x0, u, f.s[3]);
    }
    return f;
}

// The same forces as a wrench about rp.  T^T f of the natural force above is exactly (tau, F): the pseudo interpolation
// is built so that u x f_u - v x f_rd + w x f_w = tau, hence the dynamics never need the 3x3 solve.
template <typename RpOf>
__device__ __forceinline__ void external_wrench(const DevFext* __restrict__ fext, int begin, int end,
                                                const double* __restrict__ vals, V3 u, V3 v, V3 w, V3 rp, RpOf rp_of, V3& tau_sum, V3& F_sum) {
    tau_sum = mk(0.0, 0.0, 0.0), F_sum = mk(0.0, 0.0, 0.0);
    for (int e = begin; e < end; ++e) {
        V3 tau, F;
        fext_transport(fext[e], vals, u, v, w, rp, fext[e].kind == 4 ? rp_of(fext[e].other) : rp, tau, F);
        tau_sum = tau_sum + tau, F_sum = F_sum + F;
    }
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
    load9<spw>(w.S0 + 9 * w.slot_rt[idx] * spw, A);
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

// lever of a point-to-point side from the host-compiled constants (DevSide::lev, built for the model's own centre of
// mass); with per-system inertial parameters r moves by sigma (c_model - c_system)
__device__ __forceinline__ Lever lever_of_side(const WarpCtx& w, const SegCtx& c, const DevSeg& sg, const DevSide& sd) {
    Lever l;
    l.sigma = sd.lev[0];
    l.r[0] = sd.lev[1], l.r[1] = sd.lev[2], l.r[2] = sd.lev[3];
    if (w.persys) {
#pragma unroll
        for (int i = 0; i < 3; ++i) l.r[i] = fma(l.sigma, sg.kc[6 + i] - c.cn(i), l.r[i]);
    }
    return l;
}

// ---------------------------------------------------------------------------------------------
// Phase B: everything about this lane's segment that depends on the stage state only.
//   B = [u v w], reciprocal basis b_i (rows of det B^-1), Ic^-1, particular accelerations A = B^-T S with
//   S = -1/2 Bsym,  Bsym = 2 Ud^T Ud  (+ alpha Phi_r + beta Kr Qdot as symmetric matrices when stabilised:
//   natural_segment.py:429-448, :508-545, biomechanical_model.py:416-421), gyroscopic term eta, applied wrench.
// ---------------------------------------------------------------------------------------------
template <bool GENERAL, bool STAB, int S, int ES = S>
__device__ __forceinline__ void segment_setup(const WarpCtx& w, SegCtx& c, const DevSeg& sg, const DynOpts& o, int& status) {
    // ES: element stride of this lane's own state (S in the lane-minor layout, 1 when the state lies in system rows)
    const V3 rp = own3<ES>(w.Qo, 1);
    c.rp = rp;
    c.u = own3<ES>(w.Qo, 0), c.v = rp - own3<ES>(w.Qo, 2), c.w = own3<ES>(w.Qo, 3);
    const V3 b0 = cross(c.v, c.w), b1 = cross(c.w, c.u), b2 = cross(c.u, c.v);
    const double det = dot(c.u, b0);
    const double guu = dot(c.u, c.u), gvv = dot(c.v, c.v), gww = dot(c.w, c.w);
    if (!(fabs(det) > 0.0) || !isfinite(det)) status |= 1;
    else if (det * det < (kSmallPivot * kSmallPivot) * (guu * gvv * gww)) status |= 4;
    const double idet = fast_rcp(det);
    const double N[6] = {c.N3(0), c.N3(1), c.N3(2), c.N3(3), c.N3(4), c.N3(5)};
    {   // Y = B N3 B^T,  Ic = tr(Y) I - Y
        const V3 X0 = in_basis(c, N[0], N[1], N[2]), X1 = in_basis(c, N[1], N[3], N[4]), X2 = in_basis(c, N[2], N[4], N[5]);
        const double Yxx = fma(X0.x, c.u.x, fma(X1.x, c.v.x, X2.x * c.w.x));
        const double Yyy = fma(X0.y, c.u.y, fma(X1.y, c.v.y, X2.y * c.w.y));
        const double Yzz = fma(X0.z, c.u.z, fma(X1.z, c.v.z, X2.z * c.w.z));
        const double Yxy = fma(X0.x, c.u.y, fma(X1.x, c.v.y, X2.x * c.w.y));
        const double Yxz = fma(X0.x, c.u.z, fma(X1.x, c.v.z, X2.x * c.w.z));
        const double Yyz = fma(X0.y, c.u.z, fma(X1.y, c.v.z, X2.y * c.w.z));
        sym3_inverse(Yyy + Yzz, -Yxy, Yxx + Yzz, -Yxz, -Yyz, Yxx + Yyy, c.Ii, status);
    }
    const V3 ud = own3<ES>(w.Qdo, 0), vd = own3<ES>(w.Qdo, 1) - own3<ES>(w.Qdo, 2), wd = own3<ES>(w.Qdo, 3);
    double s00 = -dot(ud, ud), s01 = -dot(ud, vd), s02 = -dot(ud, wd), s11 = -dot(vd, vd), s12 = -dot(vd, wd), s22 = -dot(wd, wd);
    if (STAB && o.use_stab) {
        const double ha = 0.5 * o.alpha, hb = 0.5 * o.beta;
        s00 -= fma(ha, guu - 1.0, hb * (2.0 * dot(c.u, ud)));
        s01 -= fma(ha, dot(c.u, c.v) - sg.phic[0], hb * (dot(c.u, vd) + dot(ud, c.v)));
        s02 -= fma(ha, dot(c.u, c.w) - sg.phic[1], hb * (dot(c.u, wd) + dot(ud, c.w)));
        s11 -= fma(ha, gvv - sg.phic[2], hb * (2.0 * dot(c.v, vd)));
        s12 -= fma(ha, dot(c.v, c.w) - sg.phic[3], hb * (dot(c.v, wd) + dot(vd, c.w)));
        s22 -= fma(ha, gww - 1.0, hb * (2.0 * dot(c.w, wd)));
    }
    s00 *= idet, s01 *= idet, s02 *= idet, s11 *= idet, s12 *= idet, s22 *= idet;
    c.Au = axpy(s00, b0, axpy(s01, b1, s02 * b2));
    c.Av = axpy(s01, b0, axpy(s11, b1, s12 * b2));
    c.Aw = axpy(s02, b0, axpy(s12, b1, s22 * b2));
    {   // eta = sum_ij N3_ij B_i x A_j
        const V3 Z0 = in_accel(c, N[0], N[1], N[2]), Z1 = in_accel(c, N[1], N[3], N[4]), Z2 = in_accel(c, N[2], N[4], N[5]);
        c.tau0 = cross_acc(Z2, c.w, cross_acc(Z1, c.v, cross(Z0, c.u)));  // -eta: the factors of every cross product swapped
    }
    c.F0 = mk(0.0, 0.0, -9.81 * c.m());  // gravity acts at the centre of mass: no torque (natural_segment.py:562-572)
    if (GENERAL && w.fext != nullptr) {
        const int fb = w.ftab->begin[w.segi], fe = w.ftab->begin[w.segi + 1];
        if (fe > fb) {
            V3 te, Fe;
            const double* Qsys = w.Qsys;
            auto rp_of = [Qsys](int sgm) { return mk(Qsys[(12 * sgm + 3) * S], Qsys[(12 * sgm + 4) * S], Qsys[(12 * sgm + 5) * S]); };
            external_wrench(w.fext, fb, fe, w.fvals, c.u, c.v, c.w, rp, rp_of, te, Fe);
            const V3 cbar = in_basis(c, c.cn(0), c.cn(1), c.cn(2));
            c.F0 = c.F0 + Fe;
            c.tau0 = cross_sub(c.tau0 + te, cbar, Fe);
        }
    }
}

// One term a (x) d of a non-LIN3 row (a: 4 constant coefficients times `scale`, d: 3-vector) as it acts on the
// segment's six unknowns and on the particular acceleration: n += rho' x d, f += sigma d, kappa += d . A r.
struct Row6 {
    V3 n, f;
    double kap;
};
template <bool KAP>
__device__ __forceinline__ void row6_term(const SegCtx& c, const double* a, double scale, V3 d, Row6& r) {
    const double as[4] = {scale * a[0], scale * a[1], scale * a[2], scale * a[3]};
    const Lever lv = lever_of(c, as);
    r.n = cross_acc(in_basis(c, lv.r[0], lv.r[1], lv.r[2]), d, r.n);
    r.f = axpy(lv.sigma, d, r.f);
    if (KAP) r.kap += dot(d, in_accel(c, lv.r[0], lv.r[1], lv.r[2]));
}
// (n, fdir, kappa) of a non-LIN3 row restricted to this lane's segment; same case analysis as nonlin_k
template <bool KAP>
__device__ __forceinline__ Row6 nonlin_row6(const SegCtx& c, const DevBlk& b, int role, V3 d1, V3 d2) {
    Row6 r;
    r.n = mk(0.0, 0.0, 0.0), r.f = mk(0.0, 0.0, 0.0), r.kap = 0.0;
    const double* p = b.p;
    if (b.type == DOT) {
        row6_term<KAP>(c, role == CHILD ? p + 4 : p, 1.0, d1, r);
    } else if (b.type == CLEN) {
        row6_term<KAP>(c, role == CHILD ? p + 4 : p, role == CHILD ? -2.0 : 2.0, d1, r);
    } else if (role == CHILD) {  // SOP, blocks as coded (swapped)
        row6_term<KAP>(c, p, 1.0, d1, r);
    } else {
        row6_term<KAP>(c, p + 4, -1.0, d1, r);
        row6_term<KAP>(c, p + 8, 1.0, d2, r);
    }
    return r;
}
// the stored 6-vector of a non-LIN3 row of this lane (written in phase D)
template <int S>
__device__ __forceinline__ Row6 load_row6(const WarpCtx& w, const DevSide& sd) {
    const double* src = w.knl + sd.nl_slot * S;
    Row6 r;
    r.n = mk(src[0], src[S], src[2 * S]), r.kap = 0.0;
    r.f = (sd.ground & 2) ? mk(src[3 * S], src[4 * S], src[5 * S]) : mk(0.0, 0.0, 0.0);  // rows of directions only: f_dir = 0
    return r;
}

// Ic^-1 (rho x e_c), c = 0..2: the three columns of Ic^-1 [rho]x^T a point-to-point side needs for its pair blocks
__device__ __forceinline__ void lever_columns(const double (&di)[6], V3 rho, V3 (&V)[3]) {
    V[0] = mk(fma(di[1], rho.z, -(di[3] * rho.y)), fma(di[2], rho.z, -(di[4] * rho.y)), fma(di[4], rho.z, -(di[5] * rho.y)));
    V[1] = mk(fma(di[3], rho.x, -(di[0] * rho.z)), fma(di[4], rho.x, -(di[1] * rho.z)), fma(di[5], rho.x, -(di[3] * rho.z)));
    V[2] = mk(fma(di[0], rho.y, -(di[1] * rho.x)), fma(di[1], rho.y, -(di[2] * rho.x)), fma(di[3], rho.y, -(di[4] * rho.x)));
}

// signed coefficient vector of a LIN3 block restricted to one side: rows are  as (x) e_c
__device__ __forceinline__ void lin3_coefs(const DevBlk& b, int role, double (&as)[4]) {
    const double* a = b.p + (role == CHILD ? 4 : 0);
    const int flip = role == CHILD ? (int)0x80000000 : 0;  // sign flip on the integer pipe: the FP64 pipe is the scarce one
#pragma unroll
    for (int t = 0; t < 4; ++t) as[t] = __hiloint2double(__double2hiint(a[t]) ^ flip, __double2loint(a[t]));
}

// Pivot K of the blocked LDL^T: invert D_K, y-hat_K = D_K^-1 r_K, then for the rows I of column K that this warp
// owns (ALL: every row) the fused forward substitution and the Schur updates S_IJ -= S_IK D_K^-1 S_JK^T.
// the six unique entries of the inverse of pivot K, element i at [i * spw]
template <int spw>
__device__ __forceinline__ double* dinv_of(const WarpCtx& w, int K) {
#ifdef BIONC_SPEC_COMPACT
    // serial schedule only (one warp factors and solves): the diagonal block is dead once its pivot has been inverted, and
    // the next evaluation's assembly rewrites it
    return w.S0 + 9 * w.slot_rt[K * w.nb + K] * spw;
#else
    return w.dinv + 6 * K * spw;
#endif
}

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
        for (int i = 0; i < 6; ++i) dinv_of<spw>(w, K)[i * spw] = di[i];
        w.lamv[(3 * K) * spw] = yh.x, w.lamv[(3 * K + 1) * spw] = yh.y, w.lamv[(3 * K + 2) * spw] = yh.z;  // y-hat = D^-1 y
    }
BIONC_TAB_LOOP1
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
BIONC_TAB_LOOP1
        for (int cj = c0; cj <= ci; ++cj) {  // every J <= I of the same column: (I, J) is in the pattern
            const int J = w.colrow[cj];
            double Bj[9];
            load_block<spw>(w, J * nb + K, Bj);
            double* Sij = w.S0 + 9 * w.slot_rt[I * nb + J] * spw;
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
BIONC_TAB_LOOP1
    for (int ci = c0; ci < c1; ++ci) {
        const int I = w.colrow[ci];
        double A[9];
        load_block<spw>(w, I * w.nb + K, A);
        const V3 li = mk(w.lamv[(3 * I) * spw], w.lamv[(3 * I + 1) * spw], w.lamv[(3 * I + 2) * spw]);
        acc = axpy(li.x, mk(A[0], A[1], A[2]), axpy(li.y, mk(A[3], A[4], A[5]), axpy(li.z, mk(A[6], A[7], A[8]), acc)));
    }
    double di[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) di[i] = dinv_of<spw>(w, K)[i * spw];
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
    // experiment switches (both measured, neither kept): second solver warp; leaves / root barrier as
    // a 64-thread named barrier of the two solver warps
#ifndef BIONC_SOLVE_W1
#define BIONC_SOLVE_W1 1
#endif
#ifndef BIONC_SOLVE_NAMED
#define BIONC_SOLVE_NAMED 0
#endif
    const int K = w.segi == 0 ? 0 : (w.segi == BIONC_SOLVE_W1 ? 1 : -1);
    if (K >= 0) {
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
    if (BIONC_SOLVE_NAMED) {
        if (K >= 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (K >= 0) {
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

// lambda_r of this lane's segment from  Kr^T lambda_r = R,  R = f - Kj^T lambda_j - G Qdd  (the rows of the reference's
// KKT system that the null-space form never touches).  Kr^T lambda_r reads  R_u = 2 l0 u + l1 v + l2 w,
// R_rp = l1 u + 2 l3 v + l4 w = -R_rd,  R_w = l2 u + l4 v + 2 l5 w,  so  B^-1 [R_u R_rp R_w] = [[2 l0, l1, l2], [l1, 2 l3, l4],
// [l2, l4, 2 l5]]; the off-diagonal pairs agree up to rounding and are averaged.  Output path only: compiled into the
// forward_dynamics kernel (LAM = true), never into the fused integrator.
template <bool GENERAL, int S, int ES = S>
__device__ __forceinline__ void rigid_multipliers(const WarpCtx& w, const SegCtx& c, const DevSeg& sg, const Q12& Qdd, double* lam_r) {
    constexpr int spw = S;
    // M4 rows u, rp, w from (m, c, J), J = N3 + m c c^T   (natural_inertial_parameters.py:153-177)
    const double m = c.m(), c0 = c.cn(0), c1 = c.cn(1), c2 = c.cn(2);
    const double J00 = fma(m * c0, c0, c.N3(0)), J01 = fma(m * c0, c1, c.N3(1)), J02 = fma(m * c0, c2, c.N3(2));
    const double J11 = fma(m * c1, c1, c.N3(3)), J12 = fma(m * c1, c2, c.N3(4)), J22 = fma(m * c2, c2, c.N3(5));
    const double Mu[4] = {J00, m * c0 + J01, -J01, J02};
    const double Mr[4] = {m * c0 + J01, m + 2.0 * m * c1 + J11, -(m * c1 + J11), m * c2 + J12};
    const double Mw[4] = {J02, m * c2 + J12, -J12, J22};
    const V3 rp = ES == 1 ? c.rp : own3<ES>(w.Qo, 1);  // system rows: the state may already belong to the next tile
    const double g = -9.81 * m;
    V3 Ru = mk(0.0, 0.0, g * c0), Rr = mk(0.0, 0.0, g * (1.0 + c1)), Rw = mk(0.0, 0.0, g * c2);
    if (GENERAL && w.fext != nullptr) {
        const double* Qsys = w.Qsys;
        auto rp_of = [Qsys](int sgm) { return mk(Qsys[(12 * sgm + 3) * S], Qsys[(12 * sgm + 4) * S], Qsys[(12 * sgm + 5) * S]); };
        const Q12 fe = external_forces(w.fext, w.ftab->begin[w.segi], w.ftab->begin[w.segi + 1], w.fvals, c.u, c.v, c.w, rp, rp_of);
        Ru = Ru + fe.s[0], Rr = Rr + fe.s[1], Rw = Rw + fe.s[3];
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        Ru = axpy(-Mu[t], Qdd.s[t], Ru);
        Rr = axpy(-Mr[t], Qdd.s[t], Rr);
        Rw = axpy(-Mw[t], Qdd.s[t], Rw);
    }
    BIONC_TAB_LOOP
    for (int e = sg.side_begin; e < sg.side_end; ++e) {
        const DevSide sd = w.side[e];
        const DevBlk& b = w.blk[sd.blk];
        const double* lj = w.lamv + b.jrow * spw;
        if (!GENERAL || b.type == LIN3) {
            double as[4];
            lin3_coefs(b, sd.role, as);
            const V3 lam = mk(lj[0], lj[spw], lj[2 * spw]);
            Ru = axpy(-as[0], lam, Ru), Rr = axpy(-as[1], lam, Rr), Rw = axpy(-as[3], lam, Rw);
        } else {
            const Q12 k = rebuild_nonlin_k<S>(w, b, sd.role);
            Ru = axpy(-lj[0], k.s[0], Ru), Rr = axpy(-lj[0], k.s[1], Rr), Rw = axpy(-lj[0], k.s[3], Rw);
        }
    }
    const V3 b0 = cross(c.v, c.w), b1 = cross(c.w, c.u), b2 = cross(c.u, c.v);
    const double idet = 1.0 / dot(c.u, b0);
    lam_r[0] = 0.5 * idet * dot(b0, Ru);
    lam_r[1] = 0.5 * idet * (dot(b1, Ru) + dot(b0, Rr));
    lam_r[2] = 0.5 * idet * (dot(b2, Ru) + dot(b0, Rw));
    lam_r[3] = 0.5 * idet * dot(b1, Rr);
    lam_r[4] = 0.5 * idet * (dot(b2, Rr) + dot(b1, Rw));
    lam_r[5] = 0.5 * idet * dot(b2, Rw);
}

// ---------------------------------------------------------------------------------------------
// One forward-dynamics evaluation for the whole warp.
// Precondition: every lane has published its segment's stage state (Q, Qdot) in w.Qs / w.Qds.
// Every lane receives Qddot of its segment; lambda (holonomic order) goes to lam_out if LAM and lam_out is non-null.
// Every thread of the CTA must call this (it contains __syncthreads).
//
// GENERAL=false is the lean instantiation for models whose joints are all point-to-point (LIN3: spherical / ground
// spherical / weld halves) without external forces -- the lower-limb and full-body trees; GENERAL=true handles every
// block kind and the forces.  STAB compiles the Baumgarte terms in (they read the other segment of a joint, so the
// caller must order the publication of a stage state with a barrier when STAB or GENERAL is set).  PARKA: the particular
// accelerations wait in the thread's tensor-memory columns tmA (18 words) while the joint-level phases run.
// ---------------------------------------------------------------------------------------------
struct NoHook {
    __device__ __forceinline__ void before_assembly() const {}
    __device__ __forceinline__ void operator()() const {}
};

// hook (forward_dynamics_rows_kernel): before_assembly() is called by every thread before the first write to the joint-
// level matrix (whose area holds the previous tile's multiplier image while it is stored); operator() right after the
// first barrier that follows the assembly -- from there on no thread reads the stage state of the lean, unstabilised
// path again, so the next tile's copies may start.
template <bool GENERAL, bool STAB, bool LAM, bool PARKA, int S, bool ROWS = false, class Hook = NoHook>
__device__ __forceinline__ void eval_forward_dynamics(const WarpCtx& w, SegCtx& c, const DynOpts& o, Q12& Qdd,
                                                      double* lam_out, int& status, unsigned tmA = 0, Hook hook = Hook()) {
    constexpr int spw = S;
    const DevSeg& sg = w.seg[w.segi];
    const bool stab = STAB && o.use_stab;

    // ---- B. segment-local quantities
    constexpr int ES = ROWS ? 1 : S;  // ROWS: this lane's state lies in its system's row (lean path without stabilisation only)
    static_assert(!ROWS || (!GENERAL && !STAB), "system-row state layout: own-segment reads only");
    segment_setup<GENERAL, STAB, S, ES>(w, c, sg, o, status);
    V3 tau = c.tau0, F = c.F0;  // Newton-Euler right-hand sides, joint multipliers subtracted in phase H

    if (w.mj > 0) {
        const int nb = w.nb;
        const DevHeader* hd = w.hdr;
        // free accelerations (no joint forces); without external forces F0 is gravity: z only
        const V3 afree = sym3_apply(c.Ii, c.tau0);
        const double minv = c.minv();
        const V3 tfree = GENERAL ? minv * c.F0 : mk(0.0, 0.0, minv * c.F0.z);

        // ---- C. zero what the assembly does not overwrite (pure fill blocks, blocks of nodes made of single
        //         rows); padding rows get a unit diagonal and a zero right-hand side.  Lean models have
        //         neither, so this phase and its barrier vanish for them.
        if (hd->nzero > 0 || hd->nunit > 0) {
            BIONC_TAB_LOOP
            for (int zi = w.segi; zi < hd->nzero; zi += w.N) {
                const int z = w.zero_list[zi];
                double* dst = z >= 0 ? w.S0 + 9 * z * spw : w.S1 + 9 * (-z - 1) * spw;
#pragma unroll
                for (int e = 0; e < 9; ++e) dst[e * spw] = 0.0;
            }
            __syncthreads();
            if (w.segi == 0) {
                BIONC_TAB_LOOP
                for (int it = 0; it < hd->nunit; ++it) {
                    w.S0[w.unit_list[2 * it] * spw] = 1.0;
                    const int r = w.unit_list[2 * it + 1];
                    w.r0[r * spw] = 0.0, w.r1[r * spw] = 0.0;
                }
            }
        }

        // ---- D. right-hand side rhs' = J a_free + kappa + b_j: every segment adds its share into its own copy of the
        //         row (copy `role`, written by exactly one lane); the child side also carries the row's bias terms
        BIONC_TAB_LOOP
        for (int e = sg.side_begin; e < sg.side_end; ++e) {
            const DevSide& sd = w.side[e];
            const DevBlk& b = w.blk[sd.blk];
            double* rr = w.r0 + sd.rhs_off * spw;
            if (!GENERAL || b.type == LIN3) {
                const Lever lv = lever_of_side(w, c, sg, sd);
                const V3 rho = in_basis(c, lv.r[0], lv.r[1], lv.r[2]);
                V3 r = in_accel_acc(c, lv.r[0], lv.r[1], lv.r[2], cross(afree, rho));
                if (GENERAL) r = axpy(lv.sigma, tfree, r);
                else r.z = fma(lv.sigma, tfree.z, r.z);  // gravity only
                if (stab && sd.role == CHILD) {
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
                if (sd.ground & 1) {  // no parent lane: the child also settles the second copy
                    double* r1p = w.r1 + sd.lam_off * spw;
                    r1p[0] = 0.0, r1p[spw] = 0.0, r1p[2 * spw] = 0.0;
                }
            } else {
                V3 d1, d2;
                double bterm;
                nonlin_row<S>(w, b, sd.role, o, d1, d2, bterm);
                const Row6 r6 = nonlin_row6<true>(c, b, sd.role, d1, d2);
                double* dst = w.knl + sd.nl_slot * S;  // the row's 6-vector stays for the pair products and the back-substitution
                dst[0] = r6.n.x, dst[S] = r6.n.y, dst[2 * S] = r6.n.z;
                if (sd.ground & 2) dst[3 * S] = r6.f.x, dst[4 * S] = r6.f.y, dst[5 * S] = r6.f.z;
                rr[0] = dot(r6.n, afree) + dot(r6.f, tfree) + r6.kap + bterm;
                if (sd.ground & 1) w.r1[sd.lam_off * spw] = 0.0;
            }
        }

        if (PARKA) {  // A is not needed again before phase I
            const double a9[9] = {c.Au.x, c.Au.y, c.Au.z, c.Av.x, c.Av.y, c.Av.z, c.Aw.x, c.Aw.y, c.Aw.z};
            tmem_put(tmA, a9);
#ifdef BIONC_WAITST_EARLY
            tmem_stores_done();
#endif
        }

        // ---- E. this segment's contribution to S':  S'_ab = fdir_a . fdir_b / m + n_a . Ic^-1 n_b
        hook.before_assembly();
        BIONC_TAB_LOOP
        for (int e = sg.side_begin; e < sg.side_end; ++e) {
            const DevSide& sd = w.side[e];
            const DevBlk& b = w.blk[sd.blk];
            if (!GENERAL || b.type == LIN3) {
                // point-to-point side b against the point-to-point sides a >= b: 3x3 block gamma I + [V_c x rho_a]
                const Lever lb = lever_of_side(w, c, sg, sd);
                const V3 rho_b = in_basis(c, lb.r[0], lb.r[1], lb.r[2]);
                V3 V[3];
                lever_columns(c.Ii, rho_b, V);
                const double sbm = lb.sigma * minv;
                const int Jb = sd.node;
                BIONC_TAB_LOOP
                for (int e2 = e; e2 < sg.side_end; ++e2) {
                    const DevSide& sa = w.side[e2];
                    if (GENERAL && sa.nl_slot >= 0) continue;  // a single row: handled from its own side
                    V3 rho_a = rho_b;
                    double gamma = lb.sigma * sbm;
                    if (e2 != e) {
                        const Lever la = lever_of_side(w, c, sg, sa);
                        rho_a = in_basis(c, la.r[0], la.r[1], la.r[2]);
                        gamma = la.sigma * sbm;
                    }
                    // element (row r of a, row cc of b) = gamma delta + (rho_a x e_r) . V_cc = gamma delta + (V_cc x rho_a)_r
                    const V3 col0 = cross_acc(V[0], rho_a, mk(gamma, 0.0, 0.0)), col1 = cross_acc(V[1], rho_a, mk(0.0, gamma, 0.0)),
                             col2 = cross_acc(V[2], rho_a, mk(0.0, 0.0, gamma));
                    const int Ia = sa.node;
                    if (Ia >= Jb) {  // block (Ia, Jb): rows of a, columns of b
                        const int s2 = w.slot2_tab[Ia * nb + Jb];
                        double* Sc = (sa.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Ia * nb + Jb] * spw;
                        Sc[0 * spw] = col0.x, Sc[1 * spw] = col1.x, Sc[2 * spw] = col2.x;
                        Sc[3 * spw] = col0.y, Sc[4 * spw] = col1.y, Sc[5 * spw] = col2.y;
                        Sc[6 * spw] = col0.z, Sc[7 * spw] = col1.z, Sc[8 * spw] = col2.z;
                    } else {  // block (Jb, Ia): transposed; the copy follows this lane's role in the larger node b
                        const int s2 = w.slot2_tab[Jb * nb + Ia];
                        double* Sc = (sd.role == PARENT && s2 >= 0) ? w.S1 + 9 * s2 * spw : w.S0 + 9 * w.slot_tab[Jb * nb + Ia] * spw;
                        Sc[0 * spw] = col0.x, Sc[1 * spw] = col0.y, Sc[2 * spw] = col0.z;
                        Sc[3 * spw] = col1.x, Sc[4 * spw] = col1.y, Sc[5 * spw] = col1.z;
                        Sc[6 * spw] = col2.x, Sc[7 * spw] = col2.y, Sc[8 * spw] = col2.z;
                    }
                }
            } else {
                // single row b against every point-to-point side and against the single rows a >= b
                const Row6 rb6 = load_row6<S>(w, sd);
                const V3 yb = sym3_apply(c.Ii, rb6.n), fbm = minv * rb6.f;
                const int Ib = sd.node, rb = sd.lam_off - 3 * Ib;
                BIONC_TAB_LOOP
                for (int e2 = sg.side_begin; e2 < sg.side_end; ++e2) {
                    const DevSide& sa = w.side[e2];
                    const DevBlk& a = w.blk[sa.blk];
                    // the copy is chosen by this lane's role in the block holding the larger row
                    if (a.type == LIN3) {
                        const Lever la = lever_of_side(w, c, sg, sa);
                        const V3 rho_a = in_basis(c, la.r[0], la.r[1], la.r[2]);
                        const V3 r = cross_acc(yb, rho_a, la.sigma * fbm);
                        const int Ia = sa.node;  // LIN3 blocks are whole nodes, never Ib
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
                        const Row6 ra6 = load_row6<S>(w, sa);
                        const double acc = dot(ra6.f, fbm) + dot(ra6.n, yb);
                        const int Ia = sa.node, ra = sa.lam_off - 3 * Ia;
                        // the larger ROW decides the block orientation and whose role picks the copy
                        const bool a_larger = sa.lam_off >= sd.lam_off;
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
            hook();
            solve_two_leaves<S>(w, status);
        } else if (hd->serial_solve != 0) {
            // serial schedule (host decision: the elimination offers no parallelism across warps): the first warp
            // factors and solves everything between two barriers -- no barrier per level, no pivot inverted twice.
            __syncthreads();
            hook();
            if (w.segi == 0) {
BIONC_TAB_LOOP1
                for (int K = 0; K < nb; ++K) eliminate_pivot<true, S>(w, K, status);
BIONC_TAB_LOOP1
                for (int K = nb - 1; K >= 0; --K) back_substitute<S>(w, K);
            }
        } else {
            // level by level of the elimination tree (pivots of one level are independent: one barrier per level, the
            // first one also closes the assembly).  A block row I is always handled by warp owner[I], so concurrent
            // pivots never write the same block.
BIONC_TAB_LOOP1
            for (int lv = 0; lv < hd->nlevel; ++lv) {
                __syncthreads();
                if (lv == 0) hook();
BIONC_TAB_LOOP1
                for (int pi = w.lvlptr[lv]; pi < w.lvlptr[lv + 1]; ++pi) {
                    const int K = w.lvlpiv[pi];
                    if (((w.pivmask[K] >> w.segi) & 1) == 0) continue;  // neither the pivot's writer nor the owner of a row of column K
                    eliminate_pivot<false, S>(w, K, status);
                }
            }
            // Back substitution level by level from the roots down, the pivots of a level spread over the warps (they
            // only read lambda of higher levels); the barrier of a level also closes the y-hat / inverse pivots that the
            // factorisation stored above it.
BIONC_TAB_LOOP1
            for (int lv = hd->nlevel - 2; lv >= 0; --lv) {  // the last level holds roots only: lambda = y-hat already
                __syncthreads();
BIONC_TAB_LOOP1
                for (int pi = w.lvlptr[lv] + w.segi; pi < w.lvlptr[lv + 1]; pi += w.N) back_substitute<S>(w, w.lvlpiv[pi]);
            }
        }
        __syncthreads();

        TmemWords<18> parked;
#ifdef BIONC_WAITST_EARLY
        if (PARKA) tmem_issue(tmA, parked);  // back while phase H runs
#else
        if (PARKA) tmem_stores_done(), tmem_issue(tmA, parked);  // back while phase H runs
#endif

        // ---- H. joint forces on this segment:  F -= sum fdir lambda,  tau -= sum n lambda
        BIONC_TAB_LOOP
        for (int e = sg.side_begin; e < sg.side_end; ++e) {
            const DevSide& sd = w.side[e];
            const double* lj = w.lamv + sd.lam_off * spw;
            if (!GENERAL || sd.nl_slot < 0) {
                const Lever lv = lever_of_side(w, c, sg, sd);
                const V3 lam = mk(lj[0], lj[spw], lj[2 * spw]);
                F = axpy(-lv.sigma, lam, F);
                tau = cross_sub(tau, in_basis(c, lv.r[0], lv.r[1], lv.r[2]), lam);
                if (LAM && lam_out != nullptr && sd.role == CHILD) {
                    lam_out[sd.row_h + 0] = lam.x, lam_out[sd.row_h + 1] = lam.y, lam_out[sd.row_h + 2] = lam.z;
                }
            } else {
                const double lam = lj[0];
                const Row6 r6 = load_row6<S>(w, sd);
                F = axpy(-lam, r6.f, F);
                tau = axpy(-lam, r6.n, tau);
                if (LAM && lam_out != nullptr && sd.role == CHILD) lam_out[sd.row_h] = lam;
            }
        }
        if (PARKA) {
            double a9[9];
            tmem_wait(parked, a9);
            c.Au = mk(a9[0], a9[1], a9[2]), c.Av = mk(a9[3], a9[4], a9[5]), c.Aw = mk(a9[6], a9[7], a9[8]);
        }
    }

    // ---- I. Newton-Euler and back to natural accelerations:  Qdd = T (al, tc) + q_p
    {
        const double c0 = c.cn(0), c1 = c.cn(1), c2 = c.cn(2);
        const V3 al = sym3_apply(c.Ii, tau), tc = c.minv() * F;
        const V3 cbar = in_basis(c, c0, c1, c2);
        Qdd.s[0] = cross_acc(al, c.u, c.Au);
        Qdd.s[3] = cross_acc(al, c.w, c.Aw);
        Qdd.s[1] = in_accel_acc(c, -c0, -c1, -c2, cross_sub(tc, al, cbar));
        Qdd.s[2] = Qdd.s[1] - cross_acc(al, c.v, c.Av);
    }
    if (LAM && lam_out != nullptr) rigid_multipliers<GENERAL, S, ES>(w, c, sg, Qdd, lam_out + sg.rigid_row);
    // No barrier here: lambda_j lives in its own array, so the next evaluation may start assembling while
    // slower warps still read it; the barriers inside the next evaluation order everything else.
}

}  // namespace bionc
