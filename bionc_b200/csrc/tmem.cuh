// Tensor memory as per-thread parking space, and the CTA-size classes of the dynamics kernels.
#pragma once
#include "model.cuh"

namespace bionc {

// Kernels are instantiated per CTA-size class (threads = 32 x segments); the register budget follows from
// __launch_bounds__(MAXT, MinCtas<MAXT>): residency is what hides the latency of the dependent 3x3 chains and of the block
// barriers: 128 threads x 4 CTAs, 256 x 2 and 512 x 1 run 16 warps per SM at 128 registers per thread and the tensor-
// memory parking below takes what no longer fits the registers.  One- and two-segment models (32 / 64 threads) are
// limited by shared memory to fewer CTAs than that anyway and keep 255 registers (measured: pendulum 1.12e9 system-steps/s
// at 255 registers against 8.4e8 at 128; knee 1.9e8 with 32 systems per CTA at 255 registers against 1.2e8 with 16
// systems per CTA at 128); 96 x 4 and 384 x 1 get 168 registers, 1024 x 1 gets 64.
template <int MAXT>
struct MinCtas {
#ifdef BIONC_EXP_MINCTAS32
    static constexpr int c32 = BIONC_EXP_MINCTAS32;
#else
    static constexpr int c32 = 8;
#endif
#ifdef BIONC_EXP_MINCTAS64
    static constexpr int c64 = BIONC_EXP_MINCTAS64;
#else
    static constexpr int c64 = 4;
#endif
#ifdef BIONC_SPEC_MINCTAS
    // model-specialised build (jit.cuh): the CTAs this model's working set really leaves resident per SM -- a class's
    // register budget assumes the class's full residency, a model limited by shared memory to fewer CTAs gets the rest
    static constexpr int value = BIONC_SPEC_MINCTAS;
#else
    static constexpr int value = MAXT <= 32 ? c32 : (MAXT <= 64 ? c64 : (MAXT <= 128 ? 4 : (MAXT <= 256 ? 2 : 1)));
#endif
};

// ---- tensor memory as per-thread parking space ------------------------------------------------------
// The fused integrator carries y0 and the k-accumulator (48 doubles per thread) across every stage evaluation, and an
// evaluation carries the particular accelerations A (9 doubles) from its first phase to its last; at 128 registers
// per thread (4 CTAs per SM) none of them fits beside the evaluation's working set, and as local-memory spills they
// become L2 round trips and HBM write-backs (the first Newton-Euler kernel wrote 144 GB of evicted spills per 100-step
// launch).  The path issues no MMA, so the SM's 256 KB of tensor memory is idle: each CTA allocates columns for its
// warps and every thread parks these values in its own TMEM lane (tcgen05.st / tcgen05.ld, shape 32x32b: lane = thread
// of the warp, columns = 32-bit words).  A warp reaches the lane quarter 32 (warp % 4); warps 4.. of a CTA use the
// next group of columns.  Per-thread layout (words): y0 Q 0..23, y0 Qdot 24..47, accumulator Q 48..71, accumulator
// Qdot 72..95, A 96..113.
template <int MAXT>
struct TmemPark {
    static constexpr int groups = (MAXT + 127) / 128;
    static constexpr int words_full = 114, words_y0 = 48;
    static constexpr int budget = 512 / MinCtas<MAXT>::value;  // columns one CTA may take
    static constexpr int pow2(int need) { return need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512))); }
    // level 2: y0, accumulator and A; level 1: y0 only (the full set does not fit the CTA's share of the columns);
    // level 0: nothing (classes that run at 255 registers per thread need no parking; neither set fits)
    static constexpr int regs = 65536 / (MAXT * MinCtas<MAXT>::value);
    static constexpr int level = regs > 200 ? 0 : (pow2(words_full * groups) <= budget ? 2 : (pow2(words_y0 * groups) <= budget ? 1 : 0));
    static constexpr int words = level == 2 ? words_full : words_y0;  // per thread
    static constexpr int cols = pow2(words * groups);
};

__device__ __forceinline__ void tmem_alloc(unsigned* slot_in_smem, unsigned ncols) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(slot_in_smem);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st2(unsigned taddr, const unsigned* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_stores_done() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// N doubles <-> 2 N words.  Loads are issued first (tmem_issue) and awaited together (tmem_wait): the registers are
// in/out operands of the wait so that the compiler cannot read them before it.
template <int NW>
struct TmemWords {
    unsigned r[NW];
};
__device__ __forceinline__ void tmem_put(unsigned taddr, const double (&x)[12]) {  // 12 doubles -> 24 columns
    unsigned r[24];
#pragma unroll
    for (int i = 0; i < 12; ++i) r[2 * i] = __double2loint(x[i]), r[2 * i + 1] = __double2hiint(x[i]);
    tmem_st16(taddr, r), tmem_st8(taddr + 16, r + 16);
}
__device__ __forceinline__ void tmem_put(unsigned taddr, const double (&x)[9]) {  // 9 doubles -> 18 columns
    unsigned r[18];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[2 * i] = __double2loint(x[i]), r[2 * i + 1] = __double2hiint(x[i]);
    tmem_st16(taddr, r), tmem_st2(taddr + 16, r + 16);
}
__device__ __forceinline__ void tmem_issue(unsigned taddr, TmemWords<24>& x) {
    unsigned* r = x.r;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%24];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%25];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(taddr), "r"(taddr + 16)
        : "memory");
}
__device__ __forceinline__ void tmem_issue(unsigned taddr, TmemWords<18>& x) {
    unsigned* r = x.r;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%18];\n"
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%16, %17}, [%19];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17])
        : "r"(taddr), "r"(taddr + 16)
        : "memory");
}
__device__ __forceinline__ void tmem_wait(TmemWords<24>& x, double (&out)[12]) {
    unsigned* r = x.r;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i) out[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}
__device__ __forceinline__ void tmem_wait(TmemWords<18>& x, double (&out)[9]) {
    unsigned* r = x.r;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}

}  // namespace bionc
