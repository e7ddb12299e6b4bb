// single-evaluation kernels (see instances.cuh)
#define BIONC_DYNAMICS_ONLY
#include "kernels.cuh"
#include "instances.cuh"

namespace bionc {

const void* rk4_entry_general_nofx(int threads, int spc, bool persys);
const void* rk4_entry_general_fx(int threads, int spc, bool persys);
const void* rk4_entry_general(int threads, int spc, bool persys, bool forces) {
    return forces ? rk4_entry_general_fx(threads, spc, persys) : rk4_entry_general_nofx(threads, spc, persys);
}

const void* fd_entry(int path, int threads, int spc) {
#define X(T, S)                                                                      \
    (path == 2   ? (const void*)forward_dynamics_kernel<true, true, T, S>            \
     : path == 1 ? (const void*)forward_dynamics_kernel<false, true, T, S>           \
                 : (const void*)forward_dynamics_kernel<false, false, T, S>)
    BIONC_BY_CLASS(X)
#undef X
}

const void* fd_rows_entry(int threads) {
    if (threads <= 64) return (const void*)forward_dynamics_rows_kernel<64>;
    if (threads <= 128) return (const void*)forward_dynamics_rows_kernel<128>;
    if (threads <= 256) return (const void*)forward_dynamics_rows_kernel<256>;
    if (threads <= 512) return (const void*)forward_dynamics_rows_kernel<512>;
    return nullptr;
}

}  // namespace bionc
