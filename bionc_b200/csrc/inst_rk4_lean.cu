// fused integrator, lean paths (see instances.cuh)
#define BIONC_DYNAMICS_ONLY
#include "kernels.cuh"
#include "instances.cuh"

namespace bionc {

const void* rk4_entry_lean(bool stab, int threads, int spc, bool persys) {
#define X(T, S)                                                                                                             \
    (stab ? (persys ? (const void*)rk4_kernel<false, true, T, S, 1, false> : (const void*)rk4_kernel<false, true, T, S, 0, false>) \
          : (persys ? (const void*)rk4_kernel<false, false, T, S, 1, false> : (const void*)rk4_kernel<false, false, T, S, 0, false>))
    BIONC_BY_CLASS(X)
#undef X
}

}  // namespace bionc
