// __global__ kernels: batched forward dynamics, fused RK4, component evaluators.
#pragma once
#include "dynamics.cuh"

namespace bionc {

// What the dynamics need of the natural inertial parameters ip = [m, c(3), J00 J01 J02 J11 J12 J22]
// (natural_inertial_parameters.py:153-177): N3 = J - m c c^T, c, m, 1/m.  No inverse of M4 is involved.
__device__ __forceinline__ void inertia_to_constants(const double* __restrict__ ip, double* kc, int ks, int& status) {
    const double m = ip[0], c0 = ip[1], c1 = ip[2], c2 = ip[3];
    kc[0] = fma(-m * c0, c0, ip[4]), kc[ks] = fma(-m * c0, c1, ip[5]), kc[2 * ks] = fma(-m * c0, c2, ip[6]);
    kc[3 * ks] = fma(-m * c1, c1, ip[7]), kc[4 * ks] = fma(-m * c1, c2, ip[8]), kc[5 * ks] = fma(-m * c2, c2, ip[9]);
    kc[6 * ks] = c0, kc[7 * ks] = c1, kc[8 * ks] = c2;
    kc[9 * ks] = m, kc[10 * ks] = 1.0 / m;
    if (!(fabs(m) > 0.0) || !isfinite(m)) status |= 1;
}

// ---- per-CTA setup: copy the model blob to shared memory, carve the working set ----------------------
// FX = false: a kernel that never carries a force table (the lean path).  The compiler re-derives the working-set base
// wherever it is short of registers; without the table's pointer and size that is one constant load instead of three and
// a select (4 % of the stall samples of the lean integrator were on those loads).
// SEGI >= 0 (model-specialised builds, spec.cuh): this warp's segment as a compile-time constant.
template <int S, bool FX = true, int SEGI = -1>
__device__ __forceinline__ void setup_cta(const unsigned char* __restrict__ blob, int blob_bytes, const FextArgs& fx,
                                          unsigned char* smem, WarpCtx& w, int poison = 0, int qpad = 0) {
    const bool has_tab = FX && fx.tab != nullptr;
#ifdef BIONC_SPEC
    // The model tables are constants of this translation unit: nothing is copied, every table entry the evaluation reads
    // folds into its instruction stream.  Shared memory keeps the generic layout (the blob's area only holds the scratch
    // words of the header), so the launch geometry of the generic kernels applies unchanged.
    blob_bytes = spec::hdr.blob_bytes;
#ifdef BIONC_SPEC_COMPACT
    // compact working set (serial joint-level schedule, nothing read from a shared-memory blob): the blob's area shrinks to
    // the header's scratch words and the inverse pivots live in their dead diagonal blocks (dinv_of, dynamics.cuh); the
    // host sizes the launch accordingly (launch_rk4)
    blob_bytes = (int)((sizeof(DevHeader) + 15) / 16 * 16);
#endif
#ifdef BIONC_SPEC_ROLL_SOLVE
    {  // the joint-level factorisation keeps its loops (below): its tables are read from the shared-memory copy of the blob
        const int4* src = reinterpret_cast<const int4*>(blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = threadIdx.x; i < blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
#endif
    if (has_tab) {
        const int4* fsrc = reinterpret_cast<const int4*>(fx.tab);
        int4* fdst = reinterpret_cast<int4*>(smem + blob_bytes);
        for (int i = threadIdx.x; i < fx.tab_bytes / 16; i += blockDim.x) fdst[i] = fsrc[i];
    }
    __syncthreads();
    const DevHeader* h = &spec::hdr;
    w.hdr = h;
    w.seg = spec::seg, w.blk = spec::blk, w.side = spec::side;
    const int* sym = spec::sym;
#ifdef BIONC_SPEC_ROLL_SOLVE
    const int* symrt = reinterpret_cast<const int*>(smem + spec::hdr.off_sym);
#else
    const int* symrt = sym;
#endif
#else
    {
        const int4* src = reinterpret_cast<const int4*>(blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = threadIdx.x; i < blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
        if (has_tab) {  // this call's force table behind the blob
            const int4* fsrc = reinterpret_cast<const int4*>(fx.tab);
            int4* fdst = reinterpret_cast<int4*>(smem + blob_bytes);
            for (int i = threadIdx.x; i < fx.tab_bytes / 16; i += blockDim.x) fdst[i] = fsrc[i];
        }
    }
    __syncthreads();
    const DevHeader* h = reinterpret_cast<const DevHeader*>(smem);
    w.hdr = h;
    w.seg = reinterpret_cast<const DevSeg*>(smem + h->off_seg);
    w.blk = reinterpret_cast<const DevBlk*>(smem + h->off_blk);
    w.side = reinterpret_cast<const DevSide*>(smem + h->off_side);
    const int* sym = reinterpret_cast<const int*>(smem + h->off_sym);
    const int* symrt = sym;
#endif
    w.ftab = has_tab ? reinterpret_cast<const DevFextTab*>(smem + blob_bytes) : nullptr;
    w.fext = has_tab ? reinterpret_cast<const DevFext*>(smem + blob_bytes + sizeof(DevFextTab)) : nullptr;
    w.fvals = nullptr;
    w.N = h->N, w.mj = h->mj, w.nb = h->nb;
    w.segi = SEGI >= 0 ? SEGI : (int)(threadIdx.x >> 5);  // warp = segment
    w.lane = threadIdx.x & (S - 1);  // lane = system; with S < 32 systems per CTA the upper lanes mirror the lower ones
    double* base = reinterpret_cast<double*>(smem + blob_bytes + (has_tab ? fx.tab_bytes : 0));
    // qpad: extra doubles per system in each of the two state areas (forward_dynamics_rows_kernel keeps padded rows there)
    double* Qs = base;
    double* Qds = Qs + (12 * w.N + qpad) * S;
    w.Qsys = Qs + w.lane, w.Qdsys = Qds + w.lane;
    w.Qo = w.Qsys + 12 * S * w.segi, w.Qdo = w.Qdsys + 12 * S * w.segi;
    w.S0 = Qds + (12 * w.N + qpad) * S + w.lane;
    w.S1 = w.S0 + 9 * S * h->nslot;
    w.r0 = w.S1 + 9 * S * h->nslot2;
    w.r1 = w.r0 + 3 * S * w.nb;
    w.lamv = w.r1 + 3 * S * w.nb;
    w.dinv = w.lamv + 3 * S * w.nb;
#ifdef BIONC_SPEC_COMPACT
    const int dinv_doubles = 0;
#else
    const int dinv_doubles = 6 * S * w.nb;
#endif
    w.knl = w.dinv + dinv_doubles + S * h->rnl * w.segi;
    w.kcs = w.dinv + dinv_doubles + S * h->rnl * w.N + S * kSegConsts * w.segi;  // behind knl; present only with per-system inertia
    if (poison != 0) {
        // Debug aid: a correct evaluation writes every word of the working set before it reads it, so filling the regions
        // with NaN first must not change any result (tests/test_gpu_scale.py).  Bits: 0 S0, 1 S1, 2 r0, 3 r1, 4 lambda_j,
        // 5 inverse pivots, 6 row vectors of the non-LIN3 blocks.  Shared memory of a fresh CTA holds whatever ran before.
        const double nan = __longlong_as_double(0x7ff8dead0000beefLL);
        double* reg[7] = {w.S0 - w.lane, w.S1 - w.lane, w.r0 - w.lane, w.r1 - w.lane, w.lamv - w.lane, w.dinv - w.lane, w.dinv - w.lane + dinv_doubles};
        const int len[7] = {9 * S * h->nslot, 9 * S * h->nslot2, 3 * S * w.nb, 3 * S * w.nb, 3 * S * w.nb, dinv_doubles, S * h->rnl * w.N};
        for (int r = 0; r < 7; ++r)
            if ((poison >> r) & 1)
                for (int i = threadIdx.x; i < len[r]; i += blockDim.x) reg[r][i] = nan;
        __syncthreads();
    }
    w.slot_tab = sym;
    w.slot2_tab = w.slot_tab + w.nb * w.nb;
    w.zero_list = w.slot2_tab + w.nb * w.nb;
    w.comb_list = w.zero_list + h->nzero;
    w.unit_list = w.comb_list + 2 * h->ncomb;
    // tables of the joint-level factorisation and substitution (indexed at run time wherever their loops are kept)
    w.slot_rt = symrt;
    w.colptr = symrt + 2 * w.nb * w.nb + h->nzero + 2 * h->ncomb + 2 * h->nunit;
    w.colrow = w.colptr + w.nb + 1;
    w.lvlptr = w.colrow + h->ncol;
    w.lvlpiv = w.lvlptr + h->nlevel + 1;
    w.owner = w.lvlpiv + w.nb;
    w.pivmask = w.owner + w.nb;
    w.pivwriter = w.pivmask + w.nb;
    w.nscale = w.pivwriter + w.nb;
}

__device__ __forceinline__ void load_q12(const double* __restrict__ p, Q12& q) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2 a = p2[0], b = p2[1], c = p2[2], d = p2[3], e = p2[4], f = p2[5];
    q.s[0] = mk(a.x, a.y, b.x);
    q.s[1] = mk(b.y, c.x, c.y);
    q.s[2] = mk(d.x, d.y, e.x);
    q.s[3] = mk(e.y, f.x, f.y);
}
__device__ __forceinline__ void store_q12(double* __restrict__ p, const Q12& q) {
    double2* p2 = reinterpret_cast<double2*>(p);
    p2[0] = make_double2(q.s[0].x, q.s[0].y);
    p2[1] = make_double2(q.s[0].z, q.s[1].x);
    p2[2] = make_double2(q.s[1].y, q.s[1].z);
    p2[3] = make_double2(q.s[2].x, q.s[2].y);
    p2[4] = make_double2(q.s[2].z, q.s[3].x);
    p2[5] = make_double2(q.s[3].y, q.s[3].z);
}

// Q12 <-> tensor-memory columns (tmem.cuh works on plain arrays)
__device__ __forceinline__ void tmem_put(unsigned taddr, const Q12& q) {
    const double x[12] = {q.s[0].x, q.s[0].y, q.s[0].z, q.s[1].x, q.s[1].y, q.s[1].z, q.s[2].x, q.s[2].y, q.s[2].z, q.s[3].x, q.s[3].y, q.s[3].z};
    tmem_put(taddr, x);
}
__device__ __forceinline__ void tmem_wait(TmemWords<24>& t, Q12& q) {
    double x[12];
    tmem_wait(t, x);
#pragma unroll
    for (int k = 0; k < 4; ++k) q.s[k] = mk(x[3 * k], x[3 * k + 1], x[3 * k + 2]);
}

// PS: 0 the model's constants, 1 per-system inertial parameters, 2 decided by the call (inertia != nullptr).  The lean
// integrator is instantiated for 0 and 1: with the run-time choice the compiler re-derives the constants' base and stride
// from the kernel parameter wherever they are used (eight constant loads and selects inside the evaluation).
template <int S, int PS = 2>
__device__ __forceinline__ void load_constants(WarpCtx& w, const double* __restrict__ inertia, long long gsys,
                                               SegCtx& c, int& status) {
    const bool persys = PS == 2 ? inertia != nullptr : PS == 1;
    w.persys = persys;
    if (persys) {  // this lane's column of the CTA's constant area (only this thread ever touches it)
        inertia_to_constants(inertia + ((size_t)gsys * w.N + w.segi) * 10, w.kcs, S, status);
        c.kc = w.kcs, c.ks = S;
    } else {
        c.kc = w.seg[w.segi].kc, c.ks = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// forward_dynamics(Q, Qdot) -> Qddot, lambda       (biomechanical_model.py:351-429, batched)
// ---------------------------------------------------------------------------------------------
// The single evaluation carries no integrator state but has the multiplier output path: its register budget is set
// separately from the fused integrator's.
template <int MAXT>
struct FdMinCtas {
#ifdef BIONC_EXP_FD128
    static constexpr int value = (MAXT > 64 && MAXT <= 128) ? BIONC_EXP_FD128 : MinCtas<MAXT>::value;
#else
    static constexpr int value = MinCtas<MAXT>::value;
#endif
};

template <bool GENERAL, bool STAB, int MAXT, int S>
__global__ void __launch_bounds__(MAXT, FdMinCtas<MAXT>::value)
    forward_dynamics_kernel(const unsigned char* __restrict__ blob, int blob_bytes, FextArgs fx, const double* __restrict__ Q,
                            const double* __restrict__ Qd, const double* __restrict__ inertia, DynOpts o,
                            double* __restrict__ Qdd, double* __restrict__ lam, int* __restrict__ status_out,
                            long long B) {
    extern __shared__ __align__(16) unsigned char smem[];
    WarpCtx w;
    setup_cta<S, GENERAL>(blob, blob_bytes, fx, smem, w, o.poison);  // once per CTA: the grid is persistent, a CTA walks tiles of S systems
    const int n = 12 * w.N;
    const long long ntiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long gsys = tile * S + w.lane;
        const bool live = (int)(threadIdx.x & 31) < S && gsys < B;
        if (gsys >= B) gsys = B - 1;  // tail lanes recompute the last system, stores are masked
        int status = 0;
        SegCtx c;
        load_constants<S>(w, inertia, gsys, c, status);
        if (GENERAL && fx.vals != nullptr) w.fvals = fx.vals + (size_t)gsys * fx.nf * 6;
        {
            Q12 q, qd;
            load_q12(Q + (size_t)gsys * n + 12 * w.segi, q);
            load_q12(Qd + (size_t)gsys * n + 12 * w.segi, qd);
            // the tile this CTA takes next: into L2 while this one is evaluated (its loads then wait for L2, not for HBM)
            const long long nsys = (tile + gridDim.x) * S + w.lane;
            if (nsys < B) {
                const double* nq = Q + (size_t)nsys * n + 12 * w.segi;
                const double* nqd = Qd + (size_t)nsys * n + 12 * w.segi;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nq));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nq + 11));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nqd));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nqd + 11));
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) publish3<S>(w.Qo, t, q.s[t]), publish3<S>(w.Qdo, t, qd.s[t]);
        }
        __syncthreads();
        Q12 qdd;
        double* lam_sys = (lam != nullptr && live) ? lam + (size_t)gsys * w.hdr->m : nullptr;
        eval_forward_dynamics<GENERAL, STAB, true, false, S>(w, c, o, qdd, lam_sys, status);
        if (live) {
            store_q12(Qdd + (size_t)gsys * n + 12 * w.segi, qdd);
            if (status_out != nullptr && status != 0) atomicOr(status_out + gsys, status);
        }
        __syncthreads();  // the next tile overwrites what slower warps may still read (lambda_j, row vectors, published state)
    }
}

#ifndef BIONC_PARKA
#define BIONC_PARKA 1  // particular accelerations parked in tensor memory during the joint-level phases (experiment switch)
#endif
#ifndef BIONC_EPI_WAIT_LATE
#define BIONC_EPI_WAIT_LATE 0  // tcgen05.wait::st of the accumulator stores just before they are read again instead of right after them
#endif

// ---------------------------------------------------------------------------------------------
// RK4 with the forward-dynamics closure fused in (utils/ode_solver.py:8-53, :107-120).
// The state of a system stays on chip for all n_steps: HBM sees one read and one write of Q, Qdot
// per launch (plus the optional trajectory snapshots).  The stage state lives in shared memory
// (it has to be published there anyway for the two-segment joints); y0 is parked in tensor memory
// (TmemPark above) and the k-accumulator is a per-thread value that the compiler keeps in registers
// or spills to L2-resident local memory.  Per-system external-force values may change from step to
// step (fx.steps slices, held over the four stages of a step).  A launch integrates the steps step0 .. step0 +
// n_steps - 1 of a longer run (h_steps, traj and the force slices are indexed by the global step): the host splits a
// run in two launches around a forward_dynamics call when the multipliers of the last step are asked for.
// ---------------------------------------------------------------------------------------------
// PS: see load_constants.  FX: the call carries external forces (general path only; compiled both ways for the same
// reason as PS).
// SEGI: see setup_cta (the generic kernels run it with -1: a warp reads its segment from its index).
template <bool GENERAL, bool STAB, int MAXT, int S, int PS, bool FX, int SEGI>
__device__ __forceinline__ void rk4_body(unsigned char* smem, const unsigned char* __restrict__ blob, int blob_bytes, const FextArgs& fx,
                                         double* __restrict__ Q, double* __restrict__ Qd, const double* __restrict__ inertia, const DynOpts& o,
                                         double h_const, const double* __restrict__ h_steps, long long n_steps, long long step0,
                                         int normalize, double* __restrict__ traj, long long traj_every, int* __restrict__ status_out,
                                         long long B) {
    constexpr bool XSEG = GENERAL || STAB;  // an evaluation reads other segments' published state
    WarpCtx w;
    static_assert(GENERAL || !FX, "forces go through the general path");
    setup_cta<S, FX, SEGI>(blob, blob_bytes, fx, smem, w, o.poison);
    long long gsys = (long long)blockIdx.x * S + w.lane;
    const bool live = (int)(threadIdx.x & 31) < S && gsys < B;
    if (gsys >= B) gsys = B - 1;
    const int n = 12 * w.N;
    int status = 0;
    SegCtx c;
    load_constants<S, PS>(w, inertia, gsys, c, status);
    const double* fvals0 = (FX && fx.vals != nullptr) ? fx.vals + (size_t)gsys * fx.nf * 6 : nullptr;
    Q12 y0q, y0v;
    load_q12(Q + (size_t)gsys * n + 12 * w.segi, y0q);
    load_q12(Qd + (size_t)gsys * n + 12 * w.segi, y0v);
    constexpr int TM = TmemPark<MAXT>::level;  // 2: y0, accumulator and A in tensor memory; 1: y0; 0: nothing
    unsigned tm = 0;  // this thread's columns (layout: tmem.cuh)
    unsigned* tm_slot = reinterpret_cast<unsigned*>(&reinterpret_cast<DevHeader*>(smem)->tmem_slot);  // a free word of the header copy
    if (TM) {
        if (threadIdx.x < 32) tmem_alloc(tm_slot, TmemPark<MAXT>::cols);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm = *tm_slot + (((threadIdx.x >> 5) & 3u) << 21) + (threadIdx.x >> 7) * TmemPark<MAXT>::words;
    }

#pragma unroll 1
    for (long long step = 0; step < n_steps; ++step) {
        const double h = h_steps != nullptr ? h_steps[step0 + step] : h_const;
        if (FX && fvals0 != nullptr) w.fvals = fvals0 + (step0 + step < fx.steps ? step0 + step : fx.steps - 1) * fx.step_stride;
        Q12 accq, accv;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            accq.s[t] = mk(0.0, 0.0, 0.0), accv.s[t] = mk(0.0, 0.0, 0.0);
            publish3<S>(w.Qo, t, y0q.s[t]), publish3<S>(w.Qdo, t, y0v.s[t]);
        }
        if (TM) tmem_put(tm, y0q), tmem_put(tm + 24, y0v);
        if (TM == 2) tmem_put(tm + 48, accq), tmem_put(tm + 72, accv);  // zeros: every stage then runs the same read-add-write
        if (TM) tmem_stores_done();
        if (XSEG) __syncthreads();
        const double h6 = h / 6.0;
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            // k_stage = f(stage state): (stage velocity, Qddot)
            Q12 a;
            eval_forward_dynamics<GENERAL, STAB, false, (TM == 2) && BIONC_PARKA, S>(w, c, o, a, nullptr, status, tm + 96);
            const double wt = (stage == 0 || stage == 3) ? 1.0 : 2.0;
            const double cn = (stage == 2) ? h : 0.5 * h;  // next stage is y0 + cn * k_stage
            // other segments' published state is only read before the barriers inside the evaluation: no barrier here
            if (TM == 2) {
                // y0 and the accumulators come back from tensor memory one half at a time (Q, then Qdot: 48 registers in
                // flight); the last stage turns its half into the new y0 at once instead of writing the accumulator back
                TmemWords<24> ty, ta;
                if (BIONC_EPI_WAIT_LATE) tmem_stores_done();
                tmem_issue(tm, ty), tmem_issue(tm + 48, ta);
                tmem_wait(ty, y0q), tmem_wait(ta, accq);
                if (stage < 3) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const V3 sv = own3<S>(w.Qdo, t);
                        accq.s[t] = axpy(wt, sv, accq.s[t]);
                        publish3<S>(w.Qo, t, axpy(cn, sv, y0q.s[t]));
                    }
                    tmem_put(tm + 48, accq);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) y0q.s[t] = axpy(h6, own3<S>(w.Qdo, t) + accq.s[t], y0q.s[t]);
                }
                tmem_issue(tm + 24, ty), tmem_issue(tm + 72, ta);
                tmem_wait(ty, y0v), tmem_wait(ta, accv);
                if (stage < 3) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        accv.s[t] = axpy(wt, a.s[t], accv.s[t]);
                        publish3<S>(w.Qdo, t, axpy(cn, a.s[t], y0v.s[t]));
                    }
                    tmem_put(tm + 72, accv);
                    if (!BIONC_EPI_WAIT_LATE) tmem_stores_done();
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) y0v.s[t] = axpy(h6, a.s[t] + accv.s[t], y0v.s[t]);
                }
            } else {
                if (TM) {
                    TmemWords<24> t0, t1;
                    tmem_issue(tm, t0), tmem_issue(tm + 24, t1);
                    tmem_wait(t0, y0q), tmem_wait(t1, y0v);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const V3 sv = own3<S>(w.Qdo, t);
                    accq.s[t] = axpy(wt, sv, accq.s[t]);
                    accv.s[t] = axpy(wt, a.s[t], accv.s[t]);
                    if (stage < 3) {
                        publish3<S>(w.Qo, t, axpy(cn, sv, y0q.s[t]));
                        publish3<S>(w.Qdo, t, axpy(cn, a.s[t], y0v.s[t]));
                    }
                }
            }
            if (XSEG) __syncthreads();
        }
        if (TM < 2) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                y0q.s[t] = axpy(h6, accq.s[t], y0q.s[t]);
                y0v.s[t] = axpy(h6, accv.s[t], y0v.s[t]);
            }
        }
        if (normalize) {  // u <- u/|u|, w <- w/|w| of Q only (ode_solver.py:50-52)
            y0q.s[0] = fast_rsqrt(dot(y0q.s[0], y0q.s[0])) * y0q.s[0];
            y0q.s[3] = fast_rsqrt(dot(y0q.s[3], y0q.s[3])) * y0q.s[3];
        }
        if (traj != nullptr && ((step0 + step + 1) % traj_every) == 0 && live) {
            double* dst = traj + (((size_t)((step0 + step + 1) / traj_every - 1) * B + gsys) * 2) * n;
            store_q12(dst + 12 * w.segi, y0q);
            store_q12(dst + n + 12 * w.segi, y0v);
        }
    }
    if (TM) {  // every warp is done with its columns before warp 0 frees them
        tmem_stores_done();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) tmem_dealloc(*tm_slot, TmemPark<MAXT>::cols);
    }
    if (live) {
        double chk = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) chk += dot(y0q.s[t], y0q.s[t]) + dot(y0v.s[t], y0v.s[t]);
        if (!isfinite(chk)) status |= 2;
        store_q12(Q + (size_t)gsys * n + 12 * w.segi, y0q);
        store_q12(Qd + (size_t)gsys * n + 12 * w.segi, y0v);
        if (status_out != nullptr && status != 0) atomicOr(status_out + gsys, status);
    }
}

#ifdef BIONC_SPEC
// model-specialised build: every warp runs the copy of the integrator compiled for its segment
template <bool GENERAL, bool STAB, int MAXT, int S, int PS, bool FX, int SEGI>
struct Rk4BySegment {
    template <class... A>
    static __device__ __forceinline__ void run(unsigned warp, A&&... a) {
        if (warp == SEGI) rk4_body<GENERAL, STAB, MAXT, S, PS, FX, SEGI>(a...);
        else Rk4BySegment<GENERAL, STAB, MAXT, S, PS, FX, SEGI + 1>::run(warp, a...);
    }
};
template <bool GENERAL, bool STAB, int MAXT, int S, int PS, bool FX>
struct Rk4BySegment<GENERAL, STAB, MAXT, S, PS, FX, spec::kSegments> {
    template <class... A>
    static __device__ __forceinline__ void run(unsigned, A&&...) {}
};
#endif

template <bool GENERAL, bool STAB, int MAXT, int S, int PS, bool FX>
__global__ void __launch_bounds__(MAXT, MinCtas<MAXT>::value)
    rk4_kernel(const unsigned char* __restrict__ blob, int blob_bytes, FextArgs fx, double* __restrict__ Q, double* __restrict__ Qd,
               const double* __restrict__ inertia, DynOpts o, double h_const, const double* __restrict__ h_steps,
               long long n_steps, long long step0, int normalize, double* __restrict__ traj, long long traj_every,
               int* __restrict__ status_out, long long B) {
    extern __shared__ __align__(16) unsigned char smem[];
#ifdef BIONC_SPEC
    Rk4BySegment<GENERAL, STAB, MAXT, S, PS, FX, 0>::run(threadIdx.x >> 5, smem, blob, blob_bytes, fx, Q, Qd, inertia, o, h_const, h_steps, n_steps,
                                                         step0, normalize, traj, traj_every, status_out, B);
#else
    rk4_body<GENERAL, STAB, MAXT, S, PS, FX, -1>(smem, blob, blob_bytes, fx, Q, Qd, inertia, o, h_const, h_steps, n_steps, step0, normalize, traj,
                                                 traj_every, status_out, B);
#endif
}

// ---------------------------------------------------------------------------------------------
// Component evaluators: Phi_r, K_r, Phi_j, K_j, Phi_h, K_h, bias_h, f_ext with dense outputs.
// One thread per (system, item), item = segment (rigid rows / forces) or constraint block.
// Outputs are zero-filled by the host wrapper first; these kernels are output-bandwidth bound.
// ---------------------------------------------------------------------------------------------
struct EvalArgs {
    int which;
    int N, nblk, mj, m;
};

__device__ __forceinline__ V3 glin(const double* __restrict__ X, int s, const double* a) {
    V3 r = mk(0.0, 0.0, 0.0);
#pragma unroll
    for (int t = 0; t < 4; ++t) r = axpy(a[t], mk(X[12 * s + 3 * t], X[12 * s + 3 * t + 1], X[12 * s + 3 * t + 2]), r);
    return r;
}
// Row `row` of a Jacobian restricted to segment `seg`: the 12 entries  s1 a1 (x) d1 [+ s2 a2 (x) d2]  are WRITTEN (not
// accumulated): every structural non-zero of a system's matrix is stored exactly once, by one thread, so the tiled
// evaluator can keep a zero background in shared memory across tiles and the dense kernel needs no read-modify-write.
__device__ __forceinline__ void put_row(double* __restrict__ K, int ld, int row, int seg, const double* a, V3 d, double s) {
    double* dst = K + (size_t)row * ld + 12 * seg;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double c = s * a[t];
        dst[3 * t] = c * d.x, dst[3 * t + 1] = c * d.y, dst[3 * t + 2] = c * d.z;
    }
}
__device__ __forceinline__ void put_row2(double* __restrict__ K, int ld, int row, int seg, const double* a1, V3 d1, double s1,
                                         const double* a2, V3 d2, double s2) {
    double* dst = K + (size_t)row * ld + 12 * seg;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double c1 = s1 * a1[t], c2 = s2 * a2[t];
        dst[3 * t] = fma(c1, d1.x, c2 * d2.x), dst[3 * t + 1] = fma(c1, d1.y, c2 * d2.y), dst[3 * t + 2] = fma(c1, d1.z, c2 * d2.z);
    }
}

// One item (segment or constraint block) of one system.  `vec` / `K` point at this system's output vector / matrix
// (leading dimension ldK); X is this system's Q (or Qdot for the bias).
struct FextView {  // a call's force table as the evaluators see it (global memory)
    const DevFextTab* tab;
    const DevFext* f;
    const double* vals;  // this system's (torque, force) values or nullptr
};
__device__ __forceinline__ FextView fext_view(const FextArgs& fx, long long sys) {
    FextView v;
    v.tab = reinterpret_cast<const DevFextTab*>(fx.tab);
    v.f = fx.tab != nullptr ? reinterpret_cast<const DevFext*>(fx.tab + sizeof(DevFextTab)) : nullptr;
    v.vals = fx.vals != nullptr ? fx.vals + (size_t)sys * fx.nf * 6 : nullptr;
    return v;
}

template <bool VEC = false>  // VEC: vector quantities only (the Jacobian branches are compiled out)
__device__ __forceinline__ void eval_item(const DevSeg* __restrict__ seg, const DevBlk* __restrict__ blk,
                                          const FextView& fext, const EvalArgs& ea, int which, int item,
                                          const double* __restrict__ X, double* __restrict__ vec, double* __restrict__ K,
                                          int ldK, int sub = -1) {
    if (item < ea.N) {  // ---------------- segment items
        const int i = item;
        const DevSeg& sg = seg[i];
        const V3 u = mk(X[12 * i], X[12 * i + 1], X[12 * i + 2]);
        const V3 rp = mk(X[12 * i + 3], X[12 * i + 4], X[12 * i + 5]);
        const V3 rd = mk(X[12 * i + 6], X[12 * i + 7], X[12 * i + 8]);
        const V3 w = mk(X[12 * i + 9], X[12 * i + 10], X[12 * i + 11]);
        const V3 v = rp - rd;
        if (which == 0 || which == 4) {  // Phi_r rows
            double* o = (which == 0) ? vec + 6 * i : vec + sg.rigid_row;
            o[0] = dot(u, u) - 1.0;
            o[1] = dot(u, v) - sg.phic[0];
            o[2] = dot(u, w) - sg.phic[1];
            o[3] = dot(v, v) - sg.phic[2];
            o[4] = dot(v, w) - sg.phic[3];
            o[5] = dot(w, w) - 1.0;
        } else if (which == 6) {  // rigid bias (input is Qdot)
            double* o = vec + sg.rigid_row;
            o[0] = 2.0 * dot(u, u), o[1] = 2.0 * dot(u, v), o[2] = 2.0 * dot(u, w);
            o[3] = 2.0 * dot(v, v), o[4] = 2.0 * dot(v, w), o[5] = 2.0 * dot(w, w);
        } else if (!VEC && (which == 1 || which == 5)) {  // K_r rows
            const int r0 = (which == 1) ? 6 * i : sg.rigid_row;
            const double eu[4] = {1, 0, 0, 0}, ev[4] = {0, 1, -1, 0}, ew[4] = {0, 0, 0, 1};
            if (sub < 0 || sub == 0) put_row(K, ldK, r0 + 0, i, eu, u, 2.0);
            if (sub < 0 || sub == 1) put_row2(K, ldK, r0 + 1, i, eu, v, 1.0, ev, u, 1.0);
            if (sub < 0 || sub == 2) put_row2(K, ldK, r0 + 2, i, eu, w, 1.0, ew, u, 1.0);
            if (sub < 0 || sub == 3) put_row(K, ldK, r0 + 3, i, ev, v, 2.0);
            if (sub < 0 || sub == 4) put_row2(K, ldK, r0 + 4, i, ev, w, 1.0, ew, v, 1.0);
            if (sub < 0 || sub == 5) put_row(K, ldK, r0 + 5, i, ew, w, 2.0);
        } else if (which == 7) {  // natural external forces
            double* o = vec + 12 * i;
            Q12 f;
            if (fext.tab != nullptr) {
                auto rp_of = [X](int sgm) { return mk(X[12 * sgm + 3], X[12 * sgm + 4], X[12 * sgm + 5]); };
                f = external_forces(fext.f, fext.tab->begin[i], fext.tab->begin[i + 1], fext.vals, u, v, w, rp, rp_of);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) f.s[t] = mk(0.0, 0.0, 0.0);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) o[3 * t] = f.s[t].x, o[3 * t + 1] = f.s[t].y, o[3 * t + 2] = f.s[t].z;
        }
    } else {  // ---------------- constraint-block items
        if (which == 0 || which == 1 || which == 7) return;
        const DevBlk& b = blk[item - ea.N];
        if (sub > 2 || (sub > 0 && b.type != LIN3)) return;  // row-split callers: a block has 3 rows or 1
        const double* p = b.p;
        const bool jorder = (which == 2 || which == 3);
        const int row = jorder ? b.row_j : b.row_h;
        const bool want_phi = (which == 2 || which == 4), want_K = !VEC && (which == 3 || which == 5), want_b = (which == 6);
        double* o = vec;
        const V3 ex = mk(1, 0, 0), ey = mk(0, 1, 0), ez = mk(0, 0, 1);
        if (b.type == LIN3) {
            if (want_phi) {
                V3 phi = mk(p[8], p[9], p[10]) - glin(X, b.child, p + 4);
                if (b.parent >= 0) phi = phi + glin(X, b.parent, p);
                o[row] = phi.x, o[row + 1] = phi.y, o[row + 2] = phi.z;
            } else if (want_b) {
                o[row] = 0.0, o[row + 1] = 0.0, o[row + 2] = 0.0;  // point-to-point rows are linear: no bias
            } else if (want_K) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (sub >= 0 && sub != r) continue;
                    const V3 e = r == 0 ? ex : (r == 1 ? ey : ez);
                    put_row(K, ldK, row + r, b.child, p + 4, e, -1.0);
                    if (b.parent >= 0) put_row(K, ldK, row + r, b.parent, p, e, 1.0);
                }
            }
        } else if (b.type == DOT) {
            const V3 Dc = glin(X, b.child, p + 4);
            const V3 Dp = b.parent >= 0 ? glin(X, b.parent, p) : mk(p[0], p[1], p[2]);
            if (want_phi) o[row] = dot(Dp, Dc) - p[8];
            if (want_K) {
                put_row(K, ldK, row, b.child, p + 4, Dp, 1.0);
                if (b.parent >= 0) put_row(K, ldK, row, b.parent, p, Dc, 1.0);
            }
            if (want_b) o[row] = b.parent >= 0 ? 2.0 * dot(Dp, Dc) : 0.0;  // input is Qdot
        } else if (b.type == CLEN) {
            const V3 d = glin(X, b.parent, p) - glin(X, b.child, p + 4);
            if (want_phi) o[row] = dot(d, d) - p[8];
            if (want_K) put_row(K, ldK, row, b.parent, p, d, 2.0), put_row(K, ldK, row, b.child, p + 4, d, -2.0);
            if (want_b) o[row] = 2.0 * dot(d, d);
        } else {  // SOP
            const V3 P = glin(X, b.parent, p), A = glin(X, b.child, p + 4), nn = glin(X, b.child, p + 8);
            if (want_phi) o[row] = dot(P - A, nn) - p[12];
            if (want_K) {
                put_row2(K, ldK, row, b.parent, p + 4, nn, -1.0, p + 8, P - A, 1.0);
                put_row(K, ldK, row, b.child, p, nn, 1.0);
            }
            if (want_b) o[row] = 2.0 * dot(P, nn) - 2.0 * dot(nn, A);
        }
    }
}

// ---- 1-D bulk copies (TMA engine; SASS UBLKCP) and their mbarrier -----------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier; addresses and size are multiples of 16
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global as one bulk group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N store groups still reading their shared-memory source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// forward_dynamics for the lean path (point-to-point joints, no forces, no stabilisation, model constants): all global
// traffic goes through the bulk-copy engine instead of per-lane strided loads and stores, which kept the L1 data pipe at
// 63 % for the kernel above (profiles/r02_ncu_fd.txt) -- a lane's 96-byte slice of a 384-byte system row is 32 sectors
// per request.  Per tile of 32 systems:
//   in   one bulk copy per system row (12 N doubles) of Q and of Qdot into rows padded by 16 bytes, so that the lanes'
//        16-byte reads of their own segment fall on distinct banks; the evaluation reads its state straight from the
//        rows (ROWS layout), nothing is transposed.  The copies of the NEXT tile start inside the evaluation, right
//        after the barrier that closes the assembly (nothing reads the state after it): Q into the same rows, Qdot into
//        the other of two row buffers;
//   out  Qddot overwrites the lane's slice of the tile's Qdot rows (dead after phase B), one bulk store per row; the
//        multipliers are written as the tile's contiguous (systems, m) image into the area of the joint-level matrix
//        (dead after the back-substitution) and leave as one bulk store.  Nobody waits for the stores: the warp that
//        issued them checks that they have read their source when it starts the next copies, one evaluation later.
// Needs: joints (mj > 0), 16-byte aligned Q, Qdot, Qddot, lambda; 8 m <= 72 (nslot + nslot2) bytes per system (checked
// by the host).  A partial last tile loads the last system into the spare rows and writes its outputs with plain stores.
// ---------------------------------------------------------------------------------------------
template <int MAXT>
struct FdRowsMinCtas {
#ifdef BIONC_EXP_FDROWS128
    static constexpr int value = (MAXT > 64 && MAXT <= 128) ? BIONC_EXP_FDROWS128 : MinCtas<MAXT>::value;
#else
    static constexpr int value = (MAXT > 64 && MAXT <= 128) ? 3 : MinCtas<MAXT>::value;  // three CTAs of 74 KB per SM: 168 registers
#endif
};

struct RowsPrefetch {
    unsigned long long *bar, *img_bar;
    unsigned img_parity;
    double *qrows, *qdrows;  // destination rows of the next tile
    const double *Q, *Qd;    // this lane's source rows, nullptr when there is no next tile
    unsigned bytes;
    int ld;
    // the bulk store of the previous tile's multiplier image has read it (thread 0 arrives on img_bar when it has)
    __device__ __forceinline__ void before_assembly() const { mbar_wait(img_bar, img_parity); }
    __device__ __forceinline__ void operator()() const {
        // the copies are the business of the LAST warp: in the two-leaves and the serial schedule it has nothing to do
        // between the barriers of the joint-level solve, while the first warp is on the critical path there
        if (threadIdx.x >= blockDim.x - 32 && Q != nullptr) {
            const unsigned lane = threadIdx.x & 31;
            bulk_wait_read<0>();  // the stores of the previous tile have read the buffers written next
            if (lane == 0) mbar_expect_tx(bar, 2u * 32u * bytes);
            __syncwarp();
            bulk_load(qrows + lane * ld, Q, bytes, bar);
            bulk_load(qdrows + lane * ld, Qd, bytes, bar);
        }
    }
};

template <int MAXT>
__global__ void __launch_bounds__(MAXT, FdRowsMinCtas<MAXT>::value)
    forward_dynamics_rows_kernel(const unsigned char* __restrict__ blob, int blob_bytes, const double* __restrict__ Q,
                                 const double* __restrict__ Qd, DynOpts o, double* __restrict__ Qdd, double* __restrict__ lam,
                                 int* __restrict__ status_out, long long B) {
    constexpr int S = 32;
    extern __shared__ __align__(16) unsigned char smem[];
    WarpCtx w;
    FextArgs nofx{};
    setup_cta<S, false>(blob, blob_bytes, nofx, smem, w, o.poison, 2);
    const int n = 12 * w.N, ld = n + 2, m = w.hdr->m;
    double* Qrows = w.Qsys - w.lane;
    double* Qdrows0 = w.Qdsys - w.lane;
    double* lam_img = w.S0 - w.lane;
    // second Qdot row buffer: behind the working set (the host adds it to the launch's shared memory)
    double* Qdrows1 = w.dinv - w.lane + 6 * S * w.nb + S * w.hdr->rnl * w.N;
    unsigned long long* const barp = reinterpret_cast<unsigned long long*>(Qrows + n);  // the padding of the first row
    unsigned long long* const imgp = reinterpret_cast<unsigned long long*>(Qrows + ld + n);  // of the second
    w.Qo = Qrows + w.lane * ld + 12 * w.segi;
    w.persys = false;
    if (threadIdx.x == 0) {
        mbar_init(barp, 1), mbar_init(imgp, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (B + S - 1) / S;
    auto src_row = [&](long long tile) {  // the system this lane's row of a tile comes from
        const long long s = tile * S + w.lane;
        return s < B ? s : B - 1;
    };
    {
        RowsPrefetch first{barp, imgp, 1u, Qrows, Qdrows0, nullptr, nullptr, n * 8u, ld};
        if ((long long)blockIdx.x < ntiles) first.Q = Q + (size_t)src_row(blockIdx.x) * n, first.Qd = Qd + (size_t)src_row(blockIdx.x) * n;
        first();
    }
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long sys0 = tile * S;
        const int nlive = (int)(B - sys0 < S ? B - sys0 : S);
        double* Qdcur = (it & 1) ? Qdrows1 : Qdrows0;
        w.Qdo = Qdcur + w.lane * ld + 12 * w.segi;
        mbar_wait(barp, it & 1);
        // image barrier: phase it - 1 is completed by thread 0 after the stores of the previous tile (a fresh barrier reports
        // the phase before its first as complete)
        RowsPrefetch next{barp, imgp, (it & 1) ^ 1u, Qrows, (it & 1) ? Qdrows0 : Qdrows1, nullptr, nullptr, n * 8u, ld};
        if (tile + gridDim.x < ntiles) next.Q = Q + (size_t)src_row(tile + gridDim.x) * n, next.Qd = Qd + (size_t)src_row(tile + gridDim.x) * n;
        int status = 0;
        SegCtx c;
        c.kc = w.seg[w.segi].kc, c.ks = 1;
        Q12 qdd;
        eval_forward_dynamics<false, false, true, false, S, true, RowsPrefetch>(w, c, o, qdd, lam != nullptr ? lam_img + w.lane * m : nullptr,
                                                                               status, 0, next);
        store_q12(w.Qdo, qdd);
        if (w.lane < nlive && status_out != nullptr && status != 0) atomicOr(status_out + sys0 + w.lane, status);
        fence_async_smem();
        __syncthreads();
        if (nlive == S) {
            if (threadIdx.x >= blockDim.x - 32) {
                bulk_store(Qdd + (size_t)(sys0 + w.lane) * n, Qdcur + w.lane * ld, n * 8u);
                if (w.lane == 0) {
                    if (lam != nullptr) bulk_store(lam + (size_t)sys0 * m, lam_img, (unsigned)(S * m * 8));
                    bulk_wait_read<0>();  // only this thread waits; the others meet it at before_assembly() of the next tile
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(imgp)) : "memory");
                }
            }
        } else {  // only ever the last tile of the grid
            for (int i = threadIdx.x; i < nlive * n; i += blockDim.x) Qdd[(size_t)sys0 * n + i] = Qdcur[(i / n) * ld + i % n];
            if (lam != nullptr)
                for (int i = threadIdx.x; i < nlive * m; i += blockDim.x) lam[(size_t)sys0 * m + i] = lam_img[i];
        }
    }
    if (threadIdx.x >= blockDim.x - 32) bulk_wait_all();
}

#ifndef BIONC_DYNAMICS_ONLY  // the rest of this file is compiled into capi.cu only (see instances.cuh)

// ---------------------------------------------------------------------------------------------
// Tiled evaluators.  The outputs are dense per system (the reference's own shapes), a Jacobian mostly zeros: what the
// kernels have to do is write B x rows x cols doubles to HBM once, in full 16-byte-aligned bursts.  A CTA walks tiles;
// per tile
//   1. the states the tile needs (consecutive systems, contiguous in global memory) arrive by one bulk copy into shared
//      memory (prefetched one tile ahead, completion on an mbarrier);
//   2. the threads evaluate their units from shared memory into the tile's output image in shared memory;
//   3. one bulk copy streams the image (contiguous in global memory) out while the next tile is computed in the other
//      buffer.
// DRAM traffic = input once + output once = the algorithmic bytes; no memset, no read-modify-write on global memory.
//   eval_vec_kernel: natural external forces (f_ext); tile = T systems, unit = (system, segment).  The constraint
//                    vectors and Jacobians have their own table-driven kernels below.
// ---------------------------------------------------------------------------------------------
constexpr int kEvalThreads = 128;

struct TilePipe {  // the two-buffer pipeline shared by both kernels
    unsigned long long* bar;
    double* in_tile[2];
    double* out_tile[2];
};
__device__ __forceinline__ TilePipe tile_pipe(unsigned char* sm, size_t in_doubles, size_t out_doubles) {
    TilePipe p;
    p.bar = reinterpret_cast<unsigned long long*>(sm);
    p.in_tile[0] = reinterpret_cast<double*>(sm + 128);
    p.in_tile[1] = p.in_tile[0] + in_doubles;
    p.out_tile[0] = p.in_tile[1] + in_doubles;
    p.out_tile[1] = p.out_tile[0] + out_doubles;
    if (threadIdx.x == 0) {
        mbar_init(p.bar, 1), mbar_init(p.bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    return p;
}

__global__ void __launch_bounds__(kEvalThreads, 4)
    eval_vec_kernel(const unsigned char* __restrict__ blob, FextArgs fx, EvalArgs ea, const double* __restrict__ in,
                    double* __restrict__ out, long long B, int T, int rows) {
    extern __shared__ __align__(128) unsigned char sm[];
    const DevHeader* h = reinterpret_cast<const DevHeader*>(blob);
    const DevSeg* seg = reinterpret_cast<const DevSeg*>(blob + h->off_seg);
    const DevBlk* blk = reinterpret_cast<const DevBlk*>(blob + h->off_blk);
    const int n = 12 * ea.N, nitem = ea.N + ea.nblk, which = ea.which, tid = threadIdx.x;
    const TilePipe p = tile_pipe(sm, (size_t)T * n, (size_t)T * rows);
    const long long ntiles = (B + T - 1) / T;
    __syncthreads();
    auto tile_systems = [&](long long tile) { return (int)((B - tile * T < T) ? B - tile * T : T); };
    long long tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) {
        const unsigned bytes = (unsigned)tile_systems(tile) * n * 8u;
        mbar_expect_tx(p.bar, bytes);
        bulk_load(p.in_tile[0], in + (size_t)tile * T * n, bytes, p.bar);
    }
    for (int k = 0; tile < ntiles; tile += gridDim.x, ++k) {
        const int buf = k & 1, ns = tile_systems(tile);
        const long long next = tile + gridDim.x;
        if (tid == 0) {
            if (next < ntiles) {  // prefetch: in_tile[buf ^ 1] was last read in iteration k - 1, closed by its barrier
                const unsigned bytes = (unsigned)tile_systems(next) * n * 8u;
                mbar_expect_tx(p.bar + (buf ^ 1), bytes);
                bulk_load(p.in_tile[buf ^ 1], in + (size_t)next * T * n, bytes, p.bar + (buf ^ 1));
            }
            bulk_wait_read<1>();  // the store of iteration k - 2 has finished reading out_tile[buf]
        }
        mbar_wait(p.bar + buf, (k >> 1) & 1);
        __syncthreads();
        const double* X0 = p.in_tile[buf];
        double* O0 = p.out_tile[buf];
        for (int idx = tid; idx < ns * nitem; idx += kEvalThreads) {
            const int sys = idx / nitem, item = idx - sys * nitem;
            double* o = O0 + (size_t)sys * rows;
            eval_item<true>(seg, blk, fext_view(fx, tile * T + sys), ea, which, item, X0 + (size_t)sys * n, o, nullptr, 0);
        }
        fence_async_smem();  // generic-proxy writes of the image become visible to the bulk-copy engine
        __syncthreads();
        const size_t bytes = (size_t)ns * rows * 8;
        double* gdst = out + (size_t)tile * T * rows;
        if ((bytes & 15) == 0) {
            if (tid == 0) bulk_store(gdst, O0, (unsigned)bytes);
        } else {  // odd tail (T x rows is even for every full tile): plain stores
            for (int i = tid; i < ns * rows; i += kEvalThreads) gdst[i] = O0[i];
        }
    }
    if (tid == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Constraint vectors: Phi_r, Phi_j, Phi_h, acceleration bias.  LANE = SYSTEM, one ITEM (segment or constraint block) per
// instruction stream: a CTA takes tiles of 32 consecutive systems, stages their coordinates in shared memory with an odd
// row stride (coalesced 16-byte loads in, conflict-free reads by lane = system); warp w evaluates items w, w + 4, ... for
// its 32 systems -- the item's table entry is one broadcast read, no lane diverges, a segment's twelve coordinates serve
// its six rigid rows -- into a shared-memory image [system][row | 1], and the image (contiguous in global memory) leaves
// with coalesced stores.  (A first version evaluated every ROW as a product of two generic linear forms from a
// host-compiled table: 48 shared-memory loads per row, LSU-bound at 0.12 of the HBM peak.)
// ---------------------------------------------------------------------------------------------
constexpr int kItemThreads = 128, kItemSystems = 32;

__global__ void __launch_bounds__(kItemThreads, 8)
    items_kernel(const unsigned char* __restrict__ blob, EvalArgs ea, int rows, const double* __restrict__ in, double* __restrict__ out, long long B) {
    extern __shared__ __align__(16) double fsm[];
    const DevHeader* h = reinterpret_cast<const DevHeader*>(blob);
    const DevSeg* seg = reinterpret_cast<const DevSeg*>(blob + h->off_seg);
    const DevBlk* blk = reinterpret_cast<const DevBlk*>(blob + h->off_blk);
    const int n = 12 * ea.N, nitem = ea.N + ea.nblk;
    const int ldx = n + 1, ldo = rows | 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* Xs = fsm;                          // [32][n + 1]
    double* Os = fsm + kItemSystems * ldx;     // [32][rows | 1]
    const FextView nofx{nullptr, nullptr, nullptr};
    const long long ntiles = (B + kItemSystems - 1) / kItemSystems;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * kItemSystems;
        const int ns = (int)(B - s0 < kItemSystems ? B - s0 : kItemSystems);
        {   // ns * n contiguous doubles; n is a multiple of 12, the tile starts 16-byte aligned
            const double2* src = reinterpret_cast<const double2*>(in + (size_t)s0 * n);
            for (int i = tid; i < ns * n / 2; i += kItemThreads) {
                const double2 v = src[i];
                const int e = 2 * i, sy = e / n, k = e - sy * n;
                Xs[sy * ldx + k] = v.x, Xs[sy * ldx + k + 1] = v.y;
            }
        }
        __syncthreads();
        const double* X = Xs + (lane < ns ? lane : 0) * ldx;
        double* o = Os + lane * ldo;
        for (int item = wid; item < nitem; item += kItemThreads / 32) eval_item<true>(seg, blk, nofx, ea, ea.which, item, X, o, nullptr, 0);
        __syncthreads();
        double* dst = out + (size_t)s0 * rows;
        for (int i = tid; i < ns * rows; i += kItemThreads) {
            const int sy = i / rows, r = i - sy * rows;
            dst[i] = Os[sy * ldo + r];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Jacobians K_r, K_j, K_h.  Every entry of a constraint Jacobian of these models is an affine function of the
// coordinates (the constraints are at most quadratic), so the dense (rows x 12N) matrix of one system is a constant
// sparse affine map of its coordinate vector, compiled once per model on the host: element e = c0[e] + sum_j coef[j]
// X[off[j]], j in ptr[e] .. ptr[e+1] (most elements have no term: structural zeros).  A CTA stages KS systems'
// coordinates in shared memory; a thread owns element PAIRS (e, e + 1) and keeps KS accumulators per element, so the
// element's terms are read once per tile, and for a fixed system consecutive threads store consecutive 16-byte pairs:
// the B x rows x 12N doubles go to HBM exactly once in full coalesced bursts, with no zero fill and no scatter.
// ---------------------------------------------------------------------------------------------
constexpr int kJacThreads = 256, kJacSystems = 8;

__global__ void __launch_bounds__(kJacThreads, 4)
    jacobian_kernel(const int* __restrict__ ptr, const int* __restrict__ off, const double* __restrict__ coef, const double* __restrict__ c0,
                    int E, int n, const double* __restrict__ in, double* __restrict__ out, long long B) {
    extern __shared__ __align__(16) double jsm[];  // [KS][n]
    const int tid = threadIdx.x;
    const long long ntiles = (B + kJacSystems - 1) / kJacSystems;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * kJacSystems;
        const int ns = (int)(B - s0 < kJacSystems ? B - s0 : kJacSystems);
        {
            const double2* src = reinterpret_cast<const double2*>(in + (size_t)s0 * n);
            double2* dst = reinterpret_cast<double2*>(jsm);
            for (int i = tid; i < ns * n / 2; i += kJacThreads) dst[i] = src[i];
            for (int i = ns * n / 2 + tid; i < kJacSystems * n / 2; i += kJacThreads) dst[i] = make_double2(0.0, 0.0);  // tail tile
        }
        __syncthreads();
        for (int e = 2 * tid; e < E; e += 2 * kJacThreads) {
            double a0[kJacSystems], a1[kJacSystems];
            const double k0 = c0[e], k1 = c0[e + 1];
#pragma unroll
            for (int s = 0; s < kJacSystems; ++s) a0[s] = k0, a1[s] = k1;
            const int p0 = ptr[e], p1 = ptr[e + 1], p2 = ptr[e + 2];
            for (int j = p0; j < p1; ++j) {
                const double c = coef[j];
                const double* x = jsm + off[j];
#pragma unroll
                for (int s = 0; s < kJacSystems; ++s) a0[s] = fma(c, x[s * n], a0[s]);
            }
            for (int j = p1; j < p2; ++j) {
                const double c = coef[j];
                const double* x = jsm + off[j];
#pragma unroll
                for (int s = 0; s < kJacSystems; ++s) a1[s] = fma(c, x[s * n], a1[s]);
            }
            double* o = out + (size_t)s0 * E + e;
#pragma unroll
            for (int s = 0; s < kJacSystems; ++s)
                if (s < ns) *reinterpret_cast<double2*>(o + (size_t)s * E) = make_double2(a0[s], a1[s]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Post-processing evaluators: what the reference's users compute frame by frame after an
// integration -- marker / centre-of-mass kinematics, energies, constraint-residual norms -- on the device, so a
// trajectory does not have to go back to the host to be checked.  All of them stream Q once: HBM-bound.
// ---------------------------------------------------------------------------------------------
struct DevMarker {
    double pos[3];  // natural coordinates (n_u, n_v, n_w) of the point in its segment
    int segment, pad;
};

// natural point -> global: n_u u + (1 + n_v) rp - n_v rd + n_w w   (natural_vector.py: interpolate())
__device__ __forceinline__ V3 natural_point(const double* __restrict__ X, int seg, double nu, double nv, double nw) {
    const double* q = X + 12 * seg;
    return mk(fma(nu, q[0], fma(1.0 + nv, q[3], fma(-nv, q[6], nw * q[9]))),
              fma(nu, q[1], fma(1.0 + nv, q[4], fma(-nv, q[7], nw * q[10]))),
              fma(nu, q[2], fma(1.0 + nv, q[5], fma(-nv, q[8], nw * q[11]))));
}

// which = 8: centre-of-mass positions (biomechanical_model_markers.py:59-79) -> (B, 3, N)
// which = 9: marker positions / forward kinematics (biomechanical_model_markers.py:32-57) -> (B, 3, nb_markers)
__global__ void point_kernel(const unsigned char* __restrict__ blob, const DevMarker* __restrict__ markers, int which,
                             int N, int nitem, const double* __restrict__ Q, double* __restrict__ out, long long B) {
    const DevHeader* h = reinterpret_cast<const DevHeader*>(blob);
    const DevSeg* seg = reinterpret_cast<const DevSeg*>(blob + h->off_seg);
    const int n = 12 * N;
    const long long total = B * nitem;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long sys = idx / nitem;
        const int item = (int)(idx - sys * nitem);
        const double* X = Q + (size_t)sys * n;
        V3 p;
        if (which == 8) {
            const double* c = seg[item].mcj + 1;
            p = natural_point(X, item, c[0], c[1], c[2]);
        } else {
            const DevMarker mkr = markers[item];
            p = natural_point(X, mkr.segment, mkr.pos[0], mkr.pos[1], mkr.pos[2]);
        }
        double* o = out + (size_t)sys * 3 * nitem + item;
        o[0] = p.x, o[nitem] = p.y, o[2 * nitem] = p.z;
    }
}

// residual rows of one constraint block (the Phi branch of eval_item); returns the number of rows
__device__ __forceinline__ int block_residual(const DevBlk& b, const double* __restrict__ X, double (&phi)[3]) {
    const double* p = b.p;
    if (b.type == LIN3) {
        V3 r = mk(p[8], p[9], p[10]) - glin(X, b.child, p + 4);
        if (b.parent >= 0) r = r + glin(X, b.parent, p);
        phi[0] = r.x, phi[1] = r.y, phi[2] = r.z;
        return 3;
    }
    if (b.type == DOT) {
        const V3 Dc = glin(X, b.child, p + 4);
        const V3 Dp = b.parent >= 0 ? glin(X, b.parent, p) : mk(p[0], p[1], p[2]);
        phi[0] = dot(Dp, Dc) - p[8];
    } else if (b.type == CLEN) {
        const V3 d = glin(X, b.parent, p) - glin(X, b.child, p + 4);
        phi[0] = dot(d, d) - p[8];
    } else {
        const V3 P = glin(X, b.parent, p), A = glin(X, b.child, p + 4), nn = glin(X, b.child, p + 8);
        phi[0] = dot(P - A, nn) - p[12];
    }
    return 1;
}

// the same on the two segments held in registers (qp: parent, qc: child)
__device__ __forceinline__ V3 qlin(const Q12& q, const double* a) {
    V3 r = mk(0.0, 0.0, 0.0);
#pragma unroll
    for (int t = 0; t < 4; ++t) r = axpy(a[t], q.s[t], r);
    return r;
}
__device__ __forceinline__ int block_residual_q(const DevBlk& b, const Q12& qp, const Q12& qc, double (&phi)[3]) {
    const double* p = b.p;
    if (b.type == LIN3) {
        V3 r = mk(p[8], p[9], p[10]) - qlin(qc, p + 4);
        if (b.parent >= 0) r = r + qlin(qp, p);
        phi[0] = r.x, phi[1] = r.y, phi[2] = r.z;
        return 3;
    }
    if (b.type == DOT) {
        const V3 Dc = qlin(qc, p + 4);
        const V3 Dp = b.parent >= 0 ? qlin(qp, p) : mk(p[0], p[1], p[2]);
        phi[0] = dot(Dp, Dc) - p[8];
    } else if (b.type == CLEN) {
        const V3 d = qlin(qp, p) - qlin(qc, p + 4);
        phi[0] = dot(d, d) - p[8];
    } else {
        const V3 P = qlin(qp, p), A = qlin(qc, p + 4), nn = qlin(qc, p + 8);
        phi[0] = dot(P - A, nn) - p[12];
    }
    return 1;
}

// A group of G lanes per system (G = segments rounded up to a power of two, so a warp carries 32 / G systems
// and all its lanes load): lane-in-group = segment (then blocks in strides of G), fixed-order xor-shuffle
// reduction inside the group, so the sums are reproducible.  A group reads its system's 12N doubles contiguously.
// which = 10: kinetic energy 1/2 Qdot^T G Qdot (biomechanical_model.py:98-113)            -> (B)
// which = 11: potential energy sum_i m_i (N_com Q_i)[z] as coded, no g (natural_segment.py:775-789) -> (B)
// which = 12: norms of the rigid-body and joint residual vectors (time_series_utils.py:6-41) -> (B, 2)
template <int which>  // one instantiation per quantity: each keeps only the registers it needs
__global__ void __launch_bounds__(256)
    reduce_kernel(const unsigned char* __restrict__ blob, int N, int nblk, int G, const double* __restrict__ in,
                  double* __restrict__ out, long long B) {
    const DevHeader* h = reinterpret_cast<const DevHeader*>(blob);
    const DevSeg* seg = reinterpret_cast<const DevSeg*>(blob + h->off_seg);
    const DevBlk* blk = reinterpret_cast<const DevBlk*>(blob + h->off_blk);
    const int n = 12 * N;
    const int gl = threadIdx.x & (G - 1);  // lane inside the group
    const long long groups = ((long long)gridDim.x * blockDim.x) / G;
    const long long rounds = (B + groups - 1) / groups;  // every lane of a warp runs the same number of rounds
    long long sys = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    for (long long r = 0; r < rounds; ++r, sys += groups) {
        const bool live = sys < B;
        const double* X = in + (size_t)(live ? sys : 0) * n;
        double s0 = 0.0, s1 = 0.0;
        Q12 q;  // this lane's segment (u, rp, rd, w); zeros on idle lanes
#pragma unroll
        for (int t = 0; t < 4; ++t) q.s[t] = mk(0.0, 0.0, 0.0);
        if (live && gl < N) {
            if (which == 11) {  // the potential energy only reads the z components
                const double* x = X + 12 * gl;
                q.s[0].z = x[2], q.s[1].z = x[5], q.s[2].z = x[8], q.s[3].z = x[11];
            } else {
                load_q12(X + 12 * gl, q);  // 16-byte loads: 12 N doubles per system keep the alignment
            }
        }
        if (live && gl < N) {
            const int i = gl;
            const DevSeg& sg = seg[i];
            const V3 u = q.s[0], rp = q.s[1], rd = q.s[2], w = q.s[3];
            if (which == 10) {
                const double* ip = sg.mcj;
                const double ms = ip[0], c0 = ip[1], c1 = ip[2], c2 = ip[3];
                const double J00 = ip[4], J01 = ip[5], J02 = ip[6], J11 = ip[7], J12 = ip[8], J22 = ip[9];
                // diagonal and off-diagonal entries of M4 over the slots (u, rp, rd, w)
                s0 = 0.5 * (J00 * dot(u, u) + (ms + 2.0 * ms * c1 + J11) * dot(rp, rp) + J11 * dot(rd, rd) + J22 * dot(w, w)) +
                     (ms * c0 + J01) * dot(u, rp) - J01 * dot(u, rd) + J02 * dot(u, w) - (ms * c1 + J11) * dot(rp, rd) +
                     (ms * c2 + J12) * dot(rp, w) - J12 * dot(rd, w);
            } else if (which == 11) {
                const double* c = sg.mcj + 1;
                s0 = sg.mcj[0] * fma(c[0], u.z, fma(1.0 + c[1], rp.z, fma(-c[1], rd.z, c[2] * w.z)));
            } else {
                const V3 v = rp - rd;
                const double r0 = dot(u, u) - 1.0, r1 = dot(u, v) - sg.phic[0], r2 = dot(u, w) - sg.phic[1];
                const double r3 = dot(v, v) - sg.phic[2], r4 = dot(v, w) - sg.phic[3], r5 = dot(w, w) - 1.0;
                s0 = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3 + r4 * r4 + r5 * r5;
            }
        }
        if (which == 12) {
            // joint residuals: lane gl of the group takes blocks gl, gl + G, ...; the two segments of a block come from
            // the registers of the lanes that loaded them (shuffles inside the group), not from memory again
            const int lane0 = (threadIdx.x & 31) & ~(G - 1);
            for (int b0 = 0; b0 < nblk; b0 += G) {
                const int b = b0 + gl;
                const bool has = live && b < nblk;
                const DevBlk& bk = blk[has ? b : 0];
                const int sp = bk.parent >= 0 ? bk.parent : 0, sc = bk.child;
                Q12 qp, qc;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    qp.s[t] = mk(__shfl_sync(0xffffffffu, q.s[t].x, lane0 + sp), __shfl_sync(0xffffffffu, q.s[t].y, lane0 + sp),
                                 __shfl_sync(0xffffffffu, q.s[t].z, lane0 + sp));
                    qc.s[t] = mk(__shfl_sync(0xffffffffu, q.s[t].x, lane0 + sc), __shfl_sync(0xffffffffu, q.s[t].y, lane0 + sc),
                                 __shfl_sync(0xffffffffu, q.s[t].z, lane0 + sc));
                }
                if (has) {
                    double phi[3];
                    const int rows = block_residual_q(bk, qp, qc, phi);
                    for (int e = 0; e < rows; ++e) s1 = fma(phi[e], phi[e], s1);
                }
            }
        }
        for (int off = G >> 1; off > 0; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        }
        if (live && gl == 0) {
            if (which == 12) out[2 * sys] = sqrt(s0), out[2 * sys + 1] = sqrt(s1);
            else out[sys] = s0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pivoted fallback: the reference's own algorithm on the device (biomechanical_model.py:335-349, :423-426):
// dense augmented matrix [G K^T; K 0], right-hand side [f; -bias], LU with partial pivoting (the pivot rule
// of LAPACK dgesv: first entry of largest modulus in the column), one CTA per system, the augmented matrix in
// an L2-resident global scratch slot.  Meant for the few systems the fast path flags (SMALL_PIVOT /
// SINGULAR): indefinite mass matrices whose unpivoted LDL^T meets a small pivot although the KKT system
// itself is well conditioned.  `select` (optional) lists the systems to solve; outputs go to their rows.
// ---------------------------------------------------------------------------------------------
constexpr int kDenseThreads = 128;

__global__ void __launch_bounds__(kDenseThreads)
    dense_kkt_kernel(const unsigned char* __restrict__ blob, FextArgs fx, EvalArgs ea, const double* __restrict__ Q,
                     const double* __restrict__ Qd, const double* __restrict__ inertia, DynOpts o,
                     const long long* __restrict__ select, long long count, double* __restrict__ Qdd,
                     double* __restrict__ lam, int* __restrict__ status_out, double* __restrict__ work) {
    const DevHeader* h = reinterpret_cast<const DevHeader*>(blob);
    const DevSeg* seg = reinterpret_cast<const DevSeg*>(blob + h->off_seg);
    const DevBlk* blk = reinterpret_cast<const DevBlk*>(blob + h->off_blk);
    const int N = ea.N, n = 12 * N, m = ea.m, M = n + m, ld = M + 1, nitem = N + ea.nblk;
    const int tid = threadIdx.x, lane = tid & 31;
    double* A = work + (size_t)blockIdx.x * ((size_t)M * ld + M);  // [A | rhs], row-major
    double* tmp = A + (size_t)M * ld;                               // M doubles
    __shared__ int piv_row;
    __shared__ double piv_abs;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long sys = select != nullptr ? select[it] : it;
        const FextView fext = fext_view(fx, sys);
        const double* Xq = Q + (size_t)sys * n;
        const double* Xqd = Qd + (size_t)sys * n;
        for (int e = tid; e < M * ld + M; e += kDenseThreads) A[e] = 0.0;
        __syncthreads();
        // K (rows n.., holonomic order) and the acceleration bias
        for (int item = tid; item < nitem; item += kDenseThreads) {
            eval_item(seg, blk, fext, ea, 5, item, Xq, nullptr, A + (size_t)n * ld, ld);
            eval_item(seg, blk, fext, ea, 6, item, Xqd, tmp, nullptr, 0);
        }
        // G = blockdiag(M4_i (x) I3)  (natural_inertial_parameters.py:136-179)
        for (int e = tid; e < 16 * N; e += kDenseThreads) {
            const int i = e >> 4, a = (e >> 2) & 3, b = e & 3;
            const double* ip = inertia != nullptr ? inertia + ((size_t)sys * N + i) * 10 : seg[i].mcj;
            const double ms = ip[0], c0 = ip[1], c1 = ip[2], c2 = ip[3];
            const double J00 = ip[4], J01 = ip[5], J02 = ip[6], J11 = ip[7], J12 = ip[8], J22 = ip[9];
            const double M4[4][4] = {{J00, ms * c0 + J01, -J01, J02},
                                     {ms * c0 + J01, ms + 2.0 * ms * c1 + J11, -(ms * c1 + J11), ms * c2 + J12},
                                     {-J01, -(ms * c1 + J11), J11, -J12},
                                     {J02, ms * c2 + J12, -J12, J22}};
            double val = 0.0;
#pragma unroll
            for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
                    if (aa == a && bb == b) val = M4[aa][bb];
            for (int d = 0; d < 3; ++d) A[(size_t)(12 * i + 3 * a + d) * ld + 12 * i + 3 * b + d] = val;
        }
        __syncthreads();
        for (int r = tid; r < m; r += kDenseThreads) A[(size_t)(n + r) * ld + M] = -tmp[r];
        __syncthreads();
        if (o.use_stab) {  // bias -= alpha Phi + beta K Qdot   (biomechanical_model.py:416-421)
            for (int r = tid; r < m; r += kDenseThreads) tmp[r] = 0.0;
            __syncthreads();
            for (int item = tid; item < nitem; item += kDenseThreads) eval_item(seg, blk, fext, ea, 4, item, Xq, tmp, nullptr, 0);
            __syncthreads();
            for (int r = tid; r < m; r += kDenseThreads) {
                double kqd = 0.0;
                for (int c = 0; c < n; ++c) kqd = fma(A[(size_t)(n + r) * ld + c], Xqd[c], kqd);
                A[(size_t)(n + r) * ld + M] -= o.alpha * tmp[r] + o.beta * kqd;
            }
            __syncthreads();
        }
        // forces: gravity + natural external forces
        for (int item = tid; item < N; item += kDenseThreads) eval_item(seg, blk, fext, ea, 7, item, Xq, tmp, nullptr, 0);
        __syncthreads();
        for (int e = tid; e < 4 * N; e += kDenseThreads) {
            const int i = e >> 2, t = e & 3;
            double fg;
            if (inertia != nullptr) {
                const double* ip = inertia + ((size_t)sys * N + i) * 10;
                const double mg = ip[0] * -9.81;
                fg = t == 0 ? ip[1] * mg : (t == 1 ? (1.0 + ip[2]) * mg : (t == 2 ? -ip[2] * mg : ip[3] * mg));
            } else {
                fg = seg[i].fg[t];
            }
            for (int d = 0; d < 3; ++d) A[(size_t)(12 * i + 3 * t + d) * ld + M] = tmp[12 * i + 3 * t + d] + (d == 2 ? fg : 0.0);
        }
        for (int e = tid; e < m * n; e += kDenseThreads) {  // K^T
            const int r = e / n, c = e - r * n;
            A[(size_t)c * ld + n + r] = A[(size_t)(n + r) * ld + c];
        }
        __syncthreads();
        // LU with partial pivoting on [A | rhs]
        int st = 8;  // BIONC_STATUS_PIVOTED
        bool singular = false;
        for (int k = 0; k < M; ++k) {
            if (tid < 32) {
                double best = -1.0;
                int bi = k;
                for (int i = k + lane; i < M; i += 32) {
                    const double v = fabs(A[(size_t)i * ld + k]);
                    if (v > best) best = v, bi = i;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, best, off);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                    if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
                }
                if (lane == 0) piv_row = bi, piv_abs = best;
            }
            __syncthreads();
            const int p = piv_row;
            if (!(piv_abs > 0.0) || !isfinite(piv_abs)) {  // exact zero column: numpy raises LinAlgError here
                singular = true;
                break;
            }
            if (p != k) {
                for (int c = k + tid; c <= M; c += kDenseThreads) {
                    const double x = A[(size_t)k * ld + c];
                    A[(size_t)k * ld + c] = A[(size_t)p * ld + c];
                    A[(size_t)p * ld + c] = x;
                }
                __syncthreads();
            }
            const double akk = A[(size_t)k * ld + k];
            const int rows = M - k - 1, cols = M - k;
            for (int e = tid; e < rows * cols; e += kDenseThreads) {
                const int i = k + 1 + e / cols, j = k + 1 + e % cols;
                const double l = A[(size_t)i * ld + k] / akk;
                A[(size_t)i * ld + j] -= l * A[(size_t)k * ld + j];
            }
            __syncthreads();
        }
        if (!singular && tid < 32) {  // back substitution, x overwrites the rhs column
            for (int k = M - 1; k >= 0; --k) {
                double sum = 0.0;
                for (int j = k + 1 + lane; j < M; j += 32) sum = fma(A[(size_t)k * ld + j], A[(size_t)j * ld + M], sum);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
                if (lane == 0) A[(size_t)k * ld + M] = (A[(size_t)k * ld + M] - sum) / A[(size_t)k * ld + k];
                __syncwarp();
            }
        }
        __syncthreads();
        if (singular) st |= 1;
        bool bad = false;
        for (int c = tid; c < M; c += kDenseThreads) {
            const double x = singular ? __longlong_as_double(0x7ff8000000000000LL) : A[(size_t)c * ld + M];
            if (!isfinite(x)) bad = true;
            if (c < n) Qdd[(size_t)sys * n + c] = x;
            else if (lam != nullptr) lam[(size_t)sys * m + (c - n)] = x;
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        if (tid == 0 && status_out != nullptr) status_out[sys] = st | (any_bad && !singular ? 2 : 0);
    }
}

// A few KB of host data to device memory without any host buffer that has to outlive the call: the bytes travel in the
// kernel parameters (copied at launch time, stream-ordered, legal inside a graph capture).  Used for the per-call force
// table: cudaMemcpyAsync from pageable memory may read the source only when the stream gets there.
struct alignas(16) ParamChunk {
    unsigned char bytes[3072];
};
__global__ void param_upload_kernel(ParamChunk chunk, unsigned char* __restrict__ dst, int n) {
    for (int i = threadIdx.x; i < n / 16; i += blockDim.x) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(chunk.bytes)[i];
}

// FP64 peak probe: nothing but dependent-chain DFMAs, 8 chains per thread
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* __restrict__ out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
        x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

#endif  // BIONC_DYNAMICS_ONLY

}  // namespace bionc
