// fused integrator, general path, without external forces (see instances.cuh)
#define BIONC_DYNAMICS_ONLY
#include "kernels.cuh"
#include "instances.cuh"

namespace bionc {

const void* rk4_entry_general_nofx(int threads, int spc, bool persys) {
#define X(T, S) (persys ? (const void*)rk4_kernel<true, true, T, S, 1, false> : (const void*)rk4_kernel<true, true, T, S, 0, false>)
    BIONC_BY_CLASS(X)
#undef X
}

}  // namespace bionc
