// The dynamics kernels are instantiated per path, CTA-size class and compile-time call property -- about two hundred
// kernels -- in translation units of their own (inst_*.cu) that build.py compiles in parallel; capi.cu only sees these
// look-up functions.  Paths: 0 lean (point-to-point blocks only, no forces)   1 lean + Baumgarte   2 general (every block
// kind, forces, Baumgarte).
#pragma once

namespace bionc {

// fused integrator.  persys: per-system inertial parameters; forces: the call carries an external-force table (general
// path only).  Both are template parameters of the kernel: what is fixed at compile time is what the compiler would
// otherwise re-derive from the kernel parameters all over the evaluation (constant loads and selects in the uniform
// datapath: +11 % on the lower limb).
const void* rk4_entry_lean(bool stab, int threads, int spc, bool persys);
const void* rk4_entry_general(int threads, int spc, bool persys, bool forces);
// single evaluation with per-lane loads and stores
const void* fd_entry(int path, int threads, int spc);
// single evaluation, lean path, system rows moved by the bulk-copy engine (nullptr: no instantiation that large)
const void* fd_rows_entry(int threads);

// CTA-size classes: threads = 32 x segments rounded up to a class; 32 systems per CTA, or 16 (upper half-warps mirror the
// lower ones) for models whose working set for 32 systems would leave fewer systems resident per SM
#define BIONC_BY_CLASS(X)                      \
    if (spc == 32) {                           \
        if (threads <= 32) return X(32, 32);   \
        if (threads <= 64) return X(64, 32);   \
        if (threads <= 96) return X(96, 32);   \
        if (threads <= 128) return X(128, 32); \
        if (threads <= 256) return X(256, 32); \
        if (threads <= 384) return X(384, 32); \
        if (threads <= 512) return X(512, 32); \
        return X(1024, 32);                    \
    }                                          \
    if (threads <= 32) return X(32, 16);       \
    if (threads <= 64) return X(64, 16);       \
    if (threads <= 128) return X(128, 16);     \
    if (threads <= 256) return X(256, 16);     \
    return X(1024, 16);

}  // namespace bionc
