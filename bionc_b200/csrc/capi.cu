// C ABI (include/bionc_cuda.h): model lowering to the device blob, launch logic, host-buffer path.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bionc_cuda.h"
#include "instances.cuh"
#include "jit.cuh"
#include "kernels.cuh"

using namespace bionc;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
std::atomic<int> g_debug_poison{0};

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(BIONC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
    } while (0)

int align16(int x) { return (x + 15) & ~15; }
// the kernels read and write the per-system arrays with 16-byte vector accesses
bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

// robust inverse of a 4x4 (Gauss-Jordan, full pivoting, long double): the host has time to spare
bool invert4(const double M[4][4], double W[4][4]) {
    long double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) a[r][c] = M[r][c], a[r][4 + c] = (r == c);
    int colperm[4] = {0, 1, 2, 3};
    for (int k = 0; k < 4; ++k) {
        int pr = k, pc = k;
        long double best = 0;
        for (int r = k; r < 4; ++r)
            for (int c = k; c < 4; ++c)
                if (fabsl(a[r][c]) > best) best = fabsl(a[r][c]), pr = r, pc = c;
        if (!(best > 0)) return false;
        if (pr != k)
            for (int c = 0; c < 8; ++c) std::swap(a[k][c], a[pr][c]);
        if (pc != k) {
            for (int r = 0; r < 4; ++r) std::swap(a[r][k], a[r][pc]);
            std::swap(colperm[k], colperm[pc]);
        }
        const long double inv = 1.0L / a[k][k];
        for (int c = 0; c < 8; ++c) a[k][c] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != k) {
                const long double f = a[r][k];
                if (f != 0)
                    for (int c = 0; c < 8; ++c) a[r][c] -= f * a[k][c];
            }
    }
    // rows of the result are permuted by the column permutation
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) W[colperm[r]][c] = (double)a[r][4 + c];
    return true;
}

}  // namespace

// Launch geometry of the dynamics kernels for one model and one size of the per-call force table
struct LaunchPlan {
    int spw = 0;        // systems per CTA (32 or 16)
    size_t smem = 0;    // dynamic shared memory per CTA
    int ctas = 0;       // CTAs resident per SM (shared memory, threads, CTA-size class)
};

// Immutable after bionc_cuda_model_create: the device entry points only read it (re-entrant, any stream, any host
// thread).  Two parts are mutable, each behind its own mutex: the scratch of the host-buffer entry points (host_mu) and the
// append-only list of model-specialised kernels (jit_mu).
struct BioncModel {
    int device = 0;
    DevHeader hdr{};
    std::vector<DevSeg> seg;
    std::vector<DevBlk> blk;
    std::vector<DevSide> side;
    std::vector<int> sym;  // symbolic factorisation tables (see DevHeader::off_sym)
    int nb_markers = 0;
    DevMarker* d_markers = nullptr;  // marker table (device), outside the blob: the dynamics kernels never see it
    // device copy of the table: uploaded once, by the first call that needs the device (std::call_once), so that a
    // model can be compiled -- and its layout inspected -- on a machine without a GPU
    std::once_flag device_once;
    int device_rc = BIONC_OK;
    std::string device_err;
    unsigned char* d_blob = nullptr;
    int blob_bytes = 0;
    // host-compiled sparse affine maps of the Jacobians K_r, K_j, K_h (compile_evaluators), freed at destroy
    struct Jac {
        int* ptr = nullptr;
        int* off = nullptr;
        double* coef = nullptr;
        double* c0 = nullptr;
        int E = 0;
    } jac[3];
    int dev_smem_optin = 0, sm_smem = 0;
    LaunchPlan plan0;  // without external forces
    // model-specialised integrator kernels (bionc_cuda_model_specialize, jit.cuh): compiled and loaded on request,
    // looked up by the call shape at launch; never removed before destroy
    std::mutex jit_mu;
    std::vector<std::unique_ptr<jit::Compiled>> jit_kernels;
    // host-path scratch
    std::mutex host_mu;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // chunk pipeline of the host-buffer entry points
    double *d_Q = nullptr, *d_Qd = nullptr, *d_Qdd = nullptr, *d_lam = nullptr, *d_inertia = nullptr, *d_h = nullptr, *d_fvals = nullptr;
    int* d_status = nullptr;
    long long cap_B = 0, cap_h = 0, cap_fvals = 0;
};

namespace {

// ---------------------------------------------------------------------------------------------
// Symbolic analysis of the joint-level Schur complement S' (host, once per model).
//
// The rows of S' are grouped into 3-row "nodes": every LIN3 block is one node, the single rows of the
// other block kinds are packed three to a node (padded with unit rows).  Two nodes are coupled when
// a segment is touched by both.  The nodes are renumbered with a greedy minimum-degree ordering, the
// fill pattern of the block LDL^T is computed, and the kernel receives
//   slot_tab / slot2_tab : storage slot of block (I, J), I >= J (-1 outside the pattern); the second
//                          copy exists for blocks that can receive contributions from two segments
//   zero list            : slots the assembly does not fully overwrite (pure fill, nodes with single rows)
//   combine list         : (slot, slot2) pairs to add after the assembly
//   unit list            : diagonal elements of padding rows
//   column lists         : for every pivot node K the nodes I > K with a block (I, K)
// Sets DevBlk::jrow and the header counts; returns mj (number of real joint rows).
// ---------------------------------------------------------------------------------------------
int symbolic_analysis(BioncModel* M, int N) {
    const int nblk = (int)M->blk.size();
    // nodes in natural order: LIN3 blocks first, then singles three at a time
    std::vector<std::vector<int>> node_blocks;  // blocks of every node, in row order
    std::vector<char> mixed;                    // node holds single rows
    for (int b = 0; b < nblk; ++b)
        if (M->blk[b].type == LIN3) node_blocks.push_back({b}), mixed.push_back(0);
    int mj = 3 * (int)node_blocks.size();
    {
        std::vector<int> cur;
        for (int b = 0; b < nblk; ++b)
            if (M->blk[b].type != LIN3) {
                cur.push_back(b), ++mj;
                if (cur.size() == 3) node_blocks.push_back(cur), mixed.push_back(1), cur.clear();
            }
        if (!cur.empty()) node_blocks.push_back(cur), mixed.push_back(1);
    }
    const int nb = (int)node_blocks.size();
    // contributions: number of segments shared by two nodes
    std::vector<int> contrib(nb * nb, 0);
    for (int sgm = 0; sgm < N; ++sgm) {
        std::vector<int> touched;
        for (int I = 0; I < nb; ++I)
            for (int b : node_blocks[I])
                if (M->blk[b].child == sgm || M->blk[b].parent == sgm) {
                    touched.push_back(I);
                    break;
                }
        for (int I : touched)
            for (int J : touched) contrib[I * nb + J] += 1;
    }
    // Greedy minimum-degree ordering on the node graph (with fill), in rounds of "multiple elimination": every
    // round eliminates an independent set of minimum-degree nodes, so that leaves of different branches end up
    // on the same level of the elimination tree (fewer levels = fewer barriers in the kernel).
    std::vector<char> adj(nb * nb, 0), gone(nb, 0);
    for (int I = 0; I < nb; ++I)
        for (int J = 0; J < nb; ++J) adj[I * nb + J] = (I != J && contrib[I * nb + J] > 0);
    std::vector<int> order;
    while ((int)order.size() < nb) {
        int best_deg = 1 << 30;
        std::vector<int> deg(nb, 0);
        for (int I = 0; I < nb; ++I) {
            if (gone[I]) continue;
            for (int J = 0; J < nb; ++J) deg[I] += (!gone[J] && adj[I * nb + J]);
            if (deg[I] < best_deg) best_deg = deg[I];
        }
        std::vector<int> round;
        for (int I = 0; I < nb; ++I) {
            if (gone[I] || deg[I] != best_deg) continue;
            bool independent = true;
            for (int R : round) independent = independent && !adj[I * nb + R];
            if (independent) round.push_back(I);
        }
        for (int best : round) {
            order.push_back(best), gone[best] = 1;
            for (int I = 0; I < nb; ++I)
                for (int J = 0; J < nb; ++J)
                    if (I != J && !gone[I] && !gone[J] && adj[I * nb + best] && adj[J * nb + best]) adj[I * nb + J] = 1;
        }
    }
    std::vector<int> newidx(nb);
    for (int k = 0; k < nb; ++k) newidx[order[k]] = k;
    // renumbered structures
    std::vector<int> contrib2(nb * nb, 0);
    std::vector<char> mixed2(nb, 0), pat(nb * nb, 0);
    for (int I = 0; I < nb; ++I) {
        mixed2[newidx[I]] = mixed[I];
        for (int J = 0; J < nb; ++J) contrib2[newidx[I] * nb + newidx[J]] = contrib[I * nb + J];
    }
    for (int I = 0; I < nb; ++I) {
        int r = 0;
        for (int b : node_blocks[I]) {
            M->blk[b].jrow = 3 * newidx[I] + r;
            r += (M->blk[b].type == LIN3) ? 3 : 1;
        }
    }
    for (int I = 0; I < nb; ++I)
        for (int J = 0; J <= I; ++J) pat[I * nb + J] = (I == J) || contrib2[I * nb + J] > 0;
    for (int K = 0; K < nb; ++K)  // symbolic elimination: fill
        for (int I = K + 1; I < nb; ++I)
            if (pat[I * nb + K])
                for (int J = K + 1; J <= I; ++J)
                    if (pat[J * nb + K]) pat[I * nb + J] = 1;
    std::vector<int> slot_tab(nb * nb, -1), slot2_tab(nb * nb, -1), zero, comb, unit, colptr(nb + 1, 0), colrow;
    int nslot = 0, nslot2 = 0;
    for (int I = 0; I < nb; ++I)
        for (int J = 0; J <= I; ++J)
            if (pat[I * nb + J]) {
                const int s = nslot++;
                slot_tab[I * nb + J] = s;
                const bool any_mixed = mixed2[I] || mixed2[J];
                const int c = contrib2[I * nb + J];
                int s2 = -1;
                if (c >= 2) s2 = nslot2++, slot2_tab[I * nb + J] = s2;  // one contributing segment = one writer lane
                if (c == 0 || any_mixed) {  // not fully overwritten by the assembly
                    zero.push_back(s);
                    if (s2 >= 0) zero.push_back(-(s2 + 1));  // negative: second copy
                }
                if (s2 >= 0) comb.push_back(s), comb.push_back(s2);
            }
    // padding rows sit at the end of the node that received the last single rows (after the renumbering that node may be
    // anywhere: the kernel gets the padding rows explicitly)
    for (int I = 0; I < nb; ++I)
        if (mixed[I]) {
            int rows = (int)node_blocks[I].size();
            for (int r = rows; r < 3; ++r) {  // (diagonal element of the padding row, the row itself)
                unit.push_back(9 * slot_tab[newidx[I] * nb + newidx[I]] + 4 * r);
                unit.push_back(3 * newidx[I] + r);
            }
        }
    for (int K = 0; K < nb; ++K) {
        for (int I = K + 1; I < nb; ++I)
            if (pat[I * nb + K]) colrow.push_back(I);
        colptr[K + 1] = (int)colrow.size();
    }
    // elimination tree levels: parent(K) = first row of column K; a pivot sits one level above its deepest child.
    // Pivots of one level touch disjoint columns, so the kernel needs one barrier per level, not per pivot.
    std::vector<int> level(nb, 0), lvlptr, lvlpiv;
    for (int K = 0; K < nb; ++K)
        if (colptr[K + 1] > colptr[K]) {
            const int parent = colrow[colptr[K]];
            if (level[parent] < level[K] + 1) level[parent] = level[K] + 1;
        }
    int nlevel = 0;
    for (int K = 0; K < nb; ++K) nlevel = std::max(nlevel, level[K] + 1);
    for (int L = 0; L < nlevel; ++L) {
        lvlptr.push_back((int)lvlpiv.size());
        for (int K = 0; K < nb; ++K)
            if (level[K] == L) lvlpiv.push_back(K);
    }
    lvlptr.push_back((int)lvlpiv.size());
    M->sym.clear();
    auto append = [&](const std::vector<int>& v) { M->sym.insert(M->sym.end(), v.begin(), v.end()); };
    append(slot_tab), append(slot2_tab), append(zero), append(comb), append(unit), append(colptr), append(colrow);
    append(lvlptr), append(lvlpiv);
    // block row I is handled by warp I mod N; pivot K occupies warp 0 and the owners of the rows of its column
    // The inverse pivot and y-hat of pivot K are stored by one of the warps that needs them anyway (the owner of the
    // first row of its column; roots are spread round-robin), so no single warp has to invert every pivot of a level.
    std::vector<int> owner(nb), pivmask(nb), pivwriter(nb);
    for (int I = 0; I < nb; ++I) owner[I] = I % N;
    for (int K = 0; K < nb; ++K) {
        pivwriter[K] = colptr[K + 1] > colptr[K] ? owner[colrow[colptr[K]]] : K % N;
        unsigned mask = 1u << pivwriter[K];
        for (int ci = colptr[K]; ci < colptr[K + 1]; ++ci) mask |= 1u << owner[colrow[ci]];
        pivmask[K] = (int)mask;
    }
    append(owner), append(pivmask), append(pivwriter);
    // Schedule of the joint-level phases: level-parallel (one barrier per level, rows spread over the warps) or serial
    // on the first warp (no inner barriers, no pivot inverted twice).  Crude cost model in units of a 3x3 block
    // product: inverse 3, row (forward substitution + T) 2, pair update 2, back-substitution term 1, barrier 3.
    {
        const double c_inv = 3, c_row = 2, c_pair = 2, c_back = 1, c_bar = 3;
        double serial = 0, parallel = 0;
        for (int K = 0; K < nb; ++K) {
            const int c = colptr[K + 1] - colptr[K];
            serial += c_inv + c * c_row + 0.5 * c * (c + 1) * c_pair + (c > 0 ? 1 + c * c_back : 0);
        }
        for (int L = 0; L < nlevel; ++L) {
            std::vector<double> load(N, 0.0);
            for (int K = 0; K < nb; ++K) {
                if (level[K] != L) continue;
                for (int wi = 0; wi < N; ++wi)
                    if ((pivmask[K] >> wi) & 1) load[wi] += c_inv;
                for (int ci = colptr[K]; ci < colptr[K + 1]; ++ci) load[owner[colrow[ci]]] += c_row + (ci - colptr[K] + 1) * c_pair;
            }
            parallel += c_bar + *std::max_element(load.begin(), load.end());
        }
        for (int L = nlevel - 2; L >= 0; --L) {
            std::vector<double> load(N, 0.0);
            int k = 0;
            for (int K = 0; K < nb; ++K)
                if (level[K] == L) {
                    const int c = colptr[K + 1] - colptr[K];
                    load[k++ % N] += c > 0 ? 1 + c * c_back : 0;
                }
            parallel += c_bar + *std::max_element(load.begin(), load.end());
        }
        M->hdr.serial_solve = serial <= parallel ? 1 : 0;
        // two leaves against one root (hip / ankle against the knee), the root's diagonal and right-hand side with
        // their second copies: the leaves are eliminated by two warps at once (solve_two_leaves)
        if (nb == 3 && N >= 2 && slot_tab[1 * nb + 0] < 0 && slot_tab[2 * nb + 0] >= 0 && slot_tab[2 * nb + 1] >= 0 &&
            slot2_tab[2 * nb + 2] >= 0 && zero.empty() && unit.empty())
            M->hdr.serial_solve = 2;
    }
    M->hdr.nlevel = nlevel;
    M->hdr.nb = nb, M->hdr.nslot = nslot, M->hdr.nslot2 = nslot2;
    M->hdr.nzero = (int)zero.size(), M->hdr.ncomb = (int)comb.size() / 2, M->hdr.nunit = (int)unit.size() / 2;
    M->hdr.ncol = (int)colrow.size();
    return mj;
}


// RAII: the calling thread's current device is switched to the model's for the duration of an entry point
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define BIONC_ON_DEVICE(M)                                                                      \
    DeviceGuard guard__((M)->device);                                                           \
    if (!guard__.ok) return fail(BIONC_E_CUDA, "cudaSetDevice failed for the model's device");  \
    {                                                                                           \
        const int rc__ = ensure_device(M);                                                      \
        if (rc__ != BIONC_OK) return rc__;                                                      \
    }

// resident CTAs per SM the launch bounds of a CTA-size class are sized for (MinCtas, tmem.cuh)
int class_ctas(int threads) {
    if (threads <= 32) return MinCtas<32>::value;
    if (threads <= 64) return MinCtas<64>::value;
    if (threads <= 96) return MinCtas<96>::value;
    if (threads <= 128) return MinCtas<128>::value;
    if (threads <= 256) return MinCtas<256>::value;
    return 1;
}

// Systems per CTA and shared memory for a force table of `ftab_bytes` bytes behind the blob: 32 systems (every lane
// busy) or 16 (upper half-warps mirror), whichever keeps more systems resident per SM.
int make_plan(const BioncModel* M, int ftab_bytes, bool persys, LaunchPlan& plan) {
    const DevHeader& h = M->hdr;
    int best_resident = 0;
    plan = LaunchPlan{};
    for (int spc = 32; spc >= 16; spc >>= 1) {
        const size_t bytes = (size_t)h.blob_bytes + ftab_bytes + (size_t)cta_smem_doubles(spc, h.N, h.nb, h.nslot, h.nslot2, h.rnl, persys) * sizeof(double);
        if (bytes > (size_t)M->dev_smem_optin) continue;
        int ctas = (int)((size_t)M->sm_smem / (bytes + 1024));  // 1 KB per CTA is reserved by the driver
        const int by_threads = 2048 / (32 * h.N);
        if (ctas > by_threads) ctas = by_threads;
        const int by_class = class_ctas(32 * h.N);  // registers and tensor-memory columns of the kernel's CTA-size class
        if (ctas > by_class) ctas = by_class;
        if (ctas * spc >= best_resident && ctas > 0) best_resident = ctas * spc, plan.spw = spc, plan.smem = bytes, plan.ctas = ctas;  // ties: more CTAs, more warps
    }
    if (plan.spw == 0) return fail(BIONC_E_TOO_LARGE, "model needs more shared memory per CTA than one SM has");
    return BIONC_OK;
}

// ---------------------------------------------------------------------------------------------
// Host compilation of the Jacobian evaluator: every constraint row is  scale * sum_comp A(comp) B(comp) + c0  with two
// linear forms over the coordinates (rigid row u.v - L cos g: A = u, B = rp - rd; point-to-point row: B = 1; D_p . D_c -
// cos; |d|^2 - l^2; (P - A) . n - r), so every Jacobian entry is an affine function of the coordinates: the dense
// (rows x 12N) matrix of a system is a constant sparse affine map of its coordinate vector (jacobian_kernel).
// ---------------------------------------------------------------------------------------------
struct HostForm {
    std::vector<std::pair<int, double>> t;  // (offset of component 0, coefficient)
    bool is_const = false;
    double cst[3] = {0, 0, 0};
};
struct HostRow {
    HostForm A, B;
    bool b_one = false;
    double scale = 1.0, c0 = 0.0;
    int comp0 = 0, ncomp = 3;
    const DevBlk* sop = nullptr;  // sphere-on-plane row: its Jacobian blocks are stored as the reference codes them
};

void add_lin(HostForm& f, int seg, const double* a, double sign) {
    for (int t = 0; t < 4; ++t)
        if (a[t] != 0.0) f.t.emplace_back(12 * seg + 3 * t, sign * a[t]);
}

// row `sub` of item `item` (segment: rigid-body rows; block: joint rows); bias = acceleration-bias row (forms of Qdot)
HostRow make_row(const BioncModel* M, int item, int sub, bool bias) {
    HostRow r;
    const int N = M->hdr.N;
    if (item < N) {
        const int i = item;
        static const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {0, 1, 2, 1, 2, 2};
        auto vec = [&](int k, HostForm& f) {
            if (k == 0) f.t = {{12 * i, 1.0}};
            else if (k == 1) f.t = {{12 * i + 3, 1.0}, {12 * i + 6, -1.0}};
            else f.t = {{12 * i + 9, 1.0}};
        };
        vec(pa[sub], r.A), vec(pb[sub], r.B);
        const double* pc = M->seg[i].phic;
        const double cst[6] = {1.0, pc[0], pc[1], pc[2], pc[3], 1.0};
        r.scale = bias ? 2.0 : 1.0, r.c0 = bias ? 0.0 : -cst[sub];
        return r;
    }
    const DevBlk& b = M->blk[item - N];
    const double* p = b.p;
    if (b.type == LIN3) {
        r.b_one = true, r.comp0 = sub, r.ncomp = 1;
        if (!bias) {
            if (b.parent >= 0) add_lin(r.A, b.parent, p, 1.0);
            add_lin(r.A, b.child, p + 4, -1.0);
            r.c0 = p[8 + sub];
        }
        return r;  // bias: a zero row
    }
    r.scale = bias ? 2.0 : 1.0;
    if (b.type == DOT) {
        if (b.parent >= 0) add_lin(r.A, b.parent, p, 1.0);
        else if (!bias) r.A.is_const = true, r.A.cst[0] = p[0], r.A.cst[1] = p[1], r.A.cst[2] = p[2];
        else r.scale = 0.0;  // a ground axis does not move
        add_lin(r.B, b.child, p + 4, 1.0);
        r.c0 = bias ? 0.0 : -p[8];
    } else if (b.type == CLEN) {
        add_lin(r.A, b.parent, p, 1.0), add_lin(r.A, b.child, p + 4, -1.0);
        r.B = r.A;
        r.c0 = bias ? 0.0 : -p[8];
    } else {  // SOP
        add_lin(r.A, b.parent, p, 1.0), add_lin(r.A, b.child, p + 4, -1.0);
        add_lin(r.B, b.child, p + 8, 1.0);
        r.c0 = bias ? 0.0 : -p[12];
        r.sop = &b;
    }
    return r;
}

template <typename T>
int upload(T*& dst, const std::vector<T>& v) {
    dst = nullptr;
    if (v.empty()) return BIONC_OK;
    CUDA_TRY(cudaMalloc(&dst, v.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpy(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return BIONC_OK;
}

int compile_evaluators(BioncModel* M) {
    const int N = M->hdr.N, n = 12 * N, nblk = (int)M->blk.size(), mj = M->hdr.mj, m = M->hdr.m;
    // (item, sub) of every row in the three row orders: rigid (6N), joint-index (mj), holonomic (m)
    std::vector<std::pair<int, int>> ord[3];
    ord[0].resize(6 * N), ord[1].resize(mj), ord[2].resize(m);
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < 6; ++k) ord[0][6 * i + k] = {i, k}, ord[2][M->seg[i].rigid_row + k] = {i, k};
    for (int b = 0; b < nblk; ++b) {
        const int nr = M->blk[b].type == LIN3 ? 3 : 1;
        for (int k = 0; k < nr; ++k) ord[1][M->blk[b].row_j + k] = {N + b, k}, ord[2][M->blk[b].row_h + k] = {N + b, k};
    }
    // ---- Jacobians: element (row, col) = c0 + sum coef * X[off]
    for (int w = 0; w < 3; ++w) {
        const auto& rows = ord[w];
        const int E = (int)rows.size() * n;
        std::vector<std::vector<std::pair<int, double>>> el(E);
        std::vector<double> c0(E + 2, 0.0);
        auto add = [&](int row, int col, int off, double coef) {
            if (coef == 0.0) return;
            auto& v = el[(size_t)row * n + col];
            for (auto& t : v)
                if (t.first == off) {
                    t.second += coef;
                    return;
                }
            v.emplace_back(off, coef);
        };
        for (int r = 0; r < (int)rows.size(); ++r) {
            const HostRow hr = make_row(M, rows[r].first, rows[r].second, false);
            if (hr.sop != nullptr) {  // joints.py:676-698 as coded: the two blocks swapped between parent and child columns
                const DevBlk& b = *hr.sop;
                const double* p = b.p;
                for (int t = 0; t < 4; ++t)
                    for (int comp = 0; comp < 3; ++comp) {
                        const int colp = 12 * b.parent + 3 * t + comp, colc = 12 * b.child + 3 * t + comp;
                        for (int u = 0; u < 4; ++u) {
                            add(r, colp, 12 * b.child + 3 * u + comp, -p[4 + t] * p[8 + u]);   // -n^T N_A
                            add(r, colp, 12 * b.parent + 3 * u + comp, p[8 + t] * p[u]);        // (P - A)^T N_n
                            add(r, colp, 12 * b.child + 3 * u + comp, -p[8 + t] * p[4 + u]);
                            add(r, colc, 12 * b.child + 3 * u + comp, p[t] * p[8 + u]);         // n^T N_P
                        }
                    }
                continue;
            }
            for (int comp = hr.comp0; comp < hr.comp0 + hr.ncomp; ++comp) {
                for (auto& ta : hr.A.t) {  // d/dX[offA + comp] = scale a B(comp)
                    const int col = ta.first + comp;
                    if (hr.b_one) c0[(size_t)r * n + col] += hr.scale * ta.second;
                    else
                        for (auto& tb : hr.B.t) add(r, col, tb.first + comp, hr.scale * ta.second * tb.second);
                }
                if (!hr.b_one)
                    for (auto& tb : hr.B.t) {  // d/dX[offB + comp] = scale b A(comp)
                        const int col = tb.first + comp;
                        if (hr.A.is_const) c0[(size_t)r * n + col] += hr.scale * tb.second * hr.A.cst[comp];
                        else
                            for (auto& ta : hr.A.t) add(r, col, ta.first + comp, hr.scale * tb.second * ta.second);
                    }
            }
        }
        std::vector<int> ptr(E + 3, 0), off;
        std::vector<double> coef;
        for (int e = 0; e < E; ++e) {
            for (auto& t : el[e]) off.push_back(t.first), coef.push_back(t.second);
            ptr[e + 1] = (int)off.size();
        }
        ptr[E + 1] = ptr[E + 2] = ptr[E];
        BioncModel::Jac& J = M->jac[w];
        J.E = E;
        int rc;
        if ((rc = upload(J.ptr, ptr)) != BIONC_OK || (rc = upload(J.off, off)) != BIONC_OK || (rc = upload(J.coef, coef)) != BIONC_OK ||
            (rc = upload(J.c0, c0)) != BIONC_OK)
            return rc;
    }
    return BIONC_OK;
}

// The model blob: built and uploaded once (ensure_device).
// byte offsets of the tables inside the blob (host only)
void layout_blob(DevHeader& h, const BioncModel* M) {
    h.off_seg = align16(sizeof(DevHeader));
    h.off_blk = align16(h.off_seg + (int)(M->seg.size() * sizeof(DevSeg)));
    h.off_side = align16(h.off_blk + (int)(M->blk.size() * sizeof(DevBlk)));
    h.off_sym = align16(h.off_side + (int)(M->side.size() * sizeof(DevSide)));
    h.blob_bytes = align16(h.off_sym + (int)(M->sym.size() * sizeof(int)));
}

int build_blob(BioncModel* M) {
    DevHeader h = M->hdr;
    layout_blob(h, M);
    CUDA_TRY(cudaDeviceGetAttribute(&M->dev_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, M->device));
    CUDA_TRY(cudaDeviceGetAttribute(&M->sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, M->device));
    M->hdr = h;
    int rc = make_plan(M, 0, false, M->plan0);
    if (rc != BIONC_OK) return rc;
    h.spw = M->plan0.spw;
    M->hdr = h;
    std::vector<unsigned char> blob(h.blob_bytes, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    std::memcpy(blob.data() + h.off_seg, M->seg.data(), M->seg.size() * sizeof(DevSeg));
    if (!M->blk.empty()) std::memcpy(blob.data() + h.off_blk, M->blk.data(), M->blk.size() * sizeof(DevBlk));
    if (!M->side.empty()) std::memcpy(blob.data() + h.off_side, M->side.data(), M->side.size() * sizeof(DevSide));
    if (!M->sym.empty()) std::memcpy(blob.data() + h.off_sym, M->sym.data(), M->sym.size() * sizeof(int));
    CUDA_TRY(cudaMalloc(&M->d_blob, h.blob_bytes));
    CUDA_TRY(cudaMemcpy(M->d_blob, blob.data(), h.blob_bytes, cudaMemcpyHostToDevice));
    M->blob_bytes = h.blob_bytes;
    {
        const int rc2 = compile_evaluators(M);
        if (rc2 != BIONC_OK) return rc2;
    }
    return BIONC_OK;
}

// First use of the device by this model.  Thread-safe; every later call only reads what this leaves behind.
int ensure_device(const BioncModel* Mc) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    std::call_once(M->device_once, [M]() {
        M->device_rc = build_blob(M);
        if (M->device_rc != BIONC_OK) M->device_err = g_last_error;
    });
    if (M->device_rc != BIONC_OK) return fail(M->device_rc, M->device_err);
    return BIONC_OK;
}

// External forces of one call: the table (a few hundred bytes) is built on the host, written into a stream-ordered
// allocation on the caller's stream by a parameter-carrying kernel and freed on that stream after the launch; nothing of
// it is kept in the model and no host memory is referenced after the call returns.
struct FextCall {
    FextArgs args{};
    void* d_tab = nullptr;
    cudaStream_t st = nullptr;
    int upload(const BioncModel* M, const BioncFextDesc* fx, cudaStream_t stream) {
        st = stream;
        args = FextArgs{};
        if (fx == nullptr || fx->nb_forces <= 0) return BIONC_OK;
        const int N = M->hdr.N, nf = fx->nb_forces;
        if (fx->segment == nullptr || fx->kind == nullptr || fx->param == nullptr) return fail(BIONC_E_INVALID, "external force arrays missing");
        for (int i = 0; i < nf; ++i) {
            if (fx->segment[i] < 0 || fx->segment[i] >= N) return fail(BIONC_E_INVALID, "external force on a segment that does not exist");
            if (fx->kind[i] < 0 || fx->kind[i] > 4) return fail(BIONC_E_INVALID, "unknown external force kind");
            if (fx->kind[i] == 4) {  // the segment the force itself acts on travels in param[18]
                const double o = fx->param[(size_t)i * kFextNParam + 18];
                if (!(o >= 0 && o < N) || o != std::floor(o) || (int)o == fx->segment[i])
                    return fail(BIONC_E_INVALID, "joint reaction force without a valid other segment (param[18])");
            }
        }
        const int bytes = align16((int)(sizeof(DevFextTab) + (size_t)nf * sizeof(DevFext)));
        std::vector<unsigned char> host(bytes, 0);
        DevFextTab* tab = reinterpret_cast<DevFextTab*>(host.data());
        DevFext* f = reinterpret_cast<DevFext*>(host.data() + sizeof(DevFextTab));
        tab->nf = nf;
        int k = 0;
        for (int s = 0; s < N; ++s) {  // forces grouped by segment, insertion order kept inside a segment
            tab->begin[s] = k;
            for (int i = 0; i < nf; ++i)
                if (fx->segment[i] == s) {
                    std::memcpy(f[k].p, fx->param + (size_t)i * kFextNParam, kFextNParam * sizeof(double));
                    f[k].segment = s, f[k].kind = fx->kind[i], f[k].slot = i;
                    f[k].other = fx->kind[i] == 4 ? (int)f[k].p[18] : -1;
                    ++k;
                }
        }
        for (int s = N; s < kMaxSegments + 4; ++s) tab->begin[s] = k;
        CUDA_TRY(cudaMallocAsync(&d_tab, bytes, st));
        // the table travels in kernel parameters (copied at launch): no host buffer has to outlive this call
        for (int off = 0; off < bytes; off += (int)sizeof(ParamChunk)) {
            ParamChunk chunk;
            const int nbytes = bytes - off < (int)sizeof(ParamChunk) ? bytes - off : (int)sizeof(ParamChunk);
            std::memcpy(chunk.bytes, host.data() + off, nbytes);
            param_upload_kernel<<<1, 64, 0, st>>>(chunk, static_cast<unsigned char*>(d_tab) + off, nbytes);
            CUDA_TRY(cudaGetLastError());
        }
        args.tab = static_cast<const unsigned char*>(d_tab);
        args.tab_bytes = bytes, args.nf = nf;
        args.vals = fx->values;
        args.steps = fx->values != nullptr ? (fx->values_steps > 1 ? fx->values_steps : 1) : 0;
        args.step_stride = 0;  // set by the caller, who knows B
        return BIONC_OK;
    }
    ~FextCall() {
        if (d_tab != nullptr) cudaFreeAsync(d_tab, st);
    }
};

// the dynamics kernels of a call (instances.cuh): paths 0 lean, 1 lean + Baumgarte, 2 general
struct KernelSet {
    const void* fd;
    const void* rk4[4];  // index = 2 * (per-system inertial parameters) + (external forces; general path only)
};
KernelSet pick_kernels(int path, int threads, int spc) {
    KernelSet k{fd_entry(path, threads, spc), {nullptr, nullptr, nullptr, nullptr}};
    for (int ps = 0; ps < 2; ++ps) {
        if (path == 2) k.rk4[2 * ps] = rk4_entry_general(threads, spc, ps != 0, false), k.rk4[2 * ps + 1] = rk4_entry_general(threads, spc, ps != 0, true);
        else k.rk4[2 * ps] = rk4_entry_lean(path == 1, threads, spc, ps != 0);
    }
    return k;
}

// opt a kernel in to the device's full shared memory, once per (device, kernel)
int ensure_smem(int device, const void* kernel, int optin_bytes) {
    static std::mutex mu;
    static std::vector<std::pair<int, const void*>> done;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& d : done)
        if (d.first == device && d.second == kernel) return BIONC_OK;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin_bytes));
    // as much shared memory as the SM has: residency is limited by the per-CTA working sets
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    done.emplace_back(device, kernel);
    return BIONC_OK;
}

void dyn_options(const BioncDynOptions* opt, DynOpts& o) {
    o.use_stab = (opt != nullptr && opt->use_stabilization) ? 1 : 0;
    o.poison = g_debug_poison.load();
    o.alpha = opt != nullptr ? opt->alpha : 0.0;
    o.beta = opt != nullptr ? opt->beta : 0.0;
}

// kernel path of a call (see pick_kernels)
int kernel_path(const BioncModel* M, const DynOpts& o, const FextArgs& fx) {
    if (M->hdr.rnl > 0 || fx.tab != nullptr) return 2;
    return o.use_stab != 0 ? 1 : 0;
}

// CTA-size class of a launch (BIONC_BY_CLASS, instances.cuh)
int class_maxt(int threads, int spc) {
    if (spc == 32) {
        for (int c : {32, 64, 96, 128, 256, 384, 512})
            if (threads <= c) return c;
        return 1024;
    }
    for (int c : {32, 64, 128, 256})
        if (threads <= c) return c;
    return 1024;
}

// integrator variant of a call shape (the template arguments pick_kernels / DynCall::rk4_entry choose)
jit::Variant jit_variant(const BioncModel* M, int path, bool persys, bool forces, int spc) {
    jit::Variant v;
    v.general = path == 2, v.stab = path >= 1, v.persys = persys, v.forces = forces && path == 2;
    v.maxt = class_maxt(32 * M->hdr.N, spc), v.spc = spc;
    return v;
}

// the loaded model-specialised kernel of a variant, or nullptr (entries are never removed before destroy)
const jit::Compiled* jit_lookup(const BioncModel* Mc, const jit::Variant& v) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    std::lock_guard<std::mutex> lock(M->jit_mu);
    for (auto& k : M->jit_kernels)
        if (k->v == v && k->fn != nullptr) return k.get();
    return nullptr;
}

// Shared memory a compact specialised build gives back per CTA: the blob's area but the header's scratch words, and the
// inverse pivots (stored in their dead diagonal blocks).  Serial joint-level schedule without kept loops only.
size_t compact_savings(const BioncModel* M, int spc) {
    return (size_t)(M->hdr.blob_bytes - align16((int)sizeof(DevHeader))) + (size_t)6 * M->hdr.nb * spc * sizeof(double);
}

}  // namespace

extern "C" {

int bionc_cuda_model_create(const BioncModelDesc* d, int device, BioncModel** out) {
    if (d == nullptr || out == nullptr) return fail(BIONC_E_INVALID, "null argument");
    const int N = d->nb_segments, nblk = d->nb_blocks;
    if (N < 1 || N > kMaxSegments) return fail(BIONC_E_INVALID, "nb_segments must be in [1, 32] (one warp per segment, 1024 threads per CTA)");
    if (nblk < 0) return fail(BIONC_E_INVALID, "nb_blocks < 0");
    auto M = new BioncModel();
    M->device = device;
    M->seg.resize(N);
    M->blk.resize(nblk);
    for (int b = 0; b < nblk; ++b) {
        DevBlk& B = M->blk[b];
        std::memset(&B, 0, sizeof(B));
        B.type = d->blk_type[b], B.parent = d->blk_parent[b], B.child = d->blk_child[b];
        if (B.type < 0 || B.type > 3 || B.child < 0 || B.child >= N || B.parent < -1 || B.parent >= N || B.parent == B.child) {
            delete M;
            return fail(BIONC_E_INVALID, "malformed constraint block");
        }
        if (B.parent < 0 && (B.type == CLEN || B.type == SOP)) {
            delete M;
            return fail(BIONC_E_INVALID, "constant-length / sphere-on-plane blocks need two segments");
        }
        B.row_h = d->blk_row_h[b], B.row_j = d->blk_row_j[b];
        std::memcpy(B.p, d->blk_param + (size_t)b * kBlkNParam, kBlkNParam * sizeof(double));
    }
    const int mj = symbolic_analysis(M, N);
    int rnl = 0;
    std::vector<std::vector<double>> Winv(N);
    for (int i = 0; i < N; ++i) {
        DevSeg& S = M->seg[i];
        std::memset(&S, 0, sizeof(S));
        const double m = d->seg_mass[i];
        const double* c = d->seg_com + 3 * i;
        const double* J = d->seg_J + 9 * i;
        double M4[4][4] = {};
        M4[0][0] = J[0];
        M4[0][1] = m * c[0] + J[1];
        M4[0][2] = -J[1];
        M4[0][3] = J[2];
        M4[1][1] = m + 2 * m * c[1] + J[4];
        M4[1][2] = -(m * c[1] + J[4]);
        M4[1][3] = m * c[2] + J[5];
        M4[2][2] = J[4];
        M4[2][3] = -J[5];
        M4[3][3] = J[8];
        for (int r = 0; r < 4; ++r)
            for (int cc = 0; cc < r; ++cc) M4[r][cc] = M4[cc][r];
        if (!(std::fabs(m) > 0.0) || !std::isfinite(m)) {
            delete M;
            return fail(BIONC_E_INVALID, "segment without mass");  // reference: LinAlgError at solve time
        }
        // M4^-1 only serves the nominal scale of the joint-level nodes below; a singular M4 just disables that check
        double W[4][4];
        Winv[i].assign(10, 0.0);
        if (invert4(M4, W)) {
            int k = 0;
            for (int r = 0; r < 4; ++r)
                for (int cc = r; cc < 4; ++cc) Winv[i][k++] = 0.5 * (W[r][cc] + W[cc][r]);
        }
        // N3 = J - m c c^T (second moments about the centre of mass, natural coordinates), c, m, 1/m
        S.kc[0] = J[0] - m * c[0] * c[0], S.kc[1] = J[1] - m * c[0] * c[1], S.kc[2] = J[2] - m * c[0] * c[2];
        S.kc[3] = J[4] - m * c[1] * c[1], S.kc[4] = J[5] - m * c[1] * c[2], S.kc[5] = J[8] - m * c[2] * c[2];
        S.kc[6] = c[0], S.kc[7] = c[1], S.kc[8] = c[2], S.kc[9] = m, S.kc[10] = 1.0 / m, S.kc[11] = 0.0;
        const double g = -9.81;
        S.fg[0] = c[0] * m * g, S.fg[1] = (1 + c[1]) * m * g, S.fg[2] = -c[1] * m * g, S.fg[3] = c[2] * m * g;
        std::memcpy(S.phic, d->seg_phi_const + 4 * i, 4 * sizeof(double));
        S.mcj[0] = m, S.mcj[1] = c[0], S.mcj[2] = c[1], S.mcj[3] = c[2];
        S.mcj[4] = J[0], S.mcj[5] = J[1], S.mcj[6] = J[2], S.mcj[7] = J[4], S.mcj[8] = J[5], S.mcj[9] = J[8];
        S.rigid_row = d->seg_rigid_row[i];
        S.side_begin = (int)M->side.size();
        int nl = 0;
        for (int b = 0; b < nblk; ++b) {  // block order = holonomic order = ascending jrow
            const DevBlk& B = M->blk[b];
            if (B.child != i && B.parent != i) continue;
            DevSide sd{};
            sd.blk = b, sd.role = (B.child == i) ? CHILD : PARENT;
            sd.node = B.jrow / 3, sd.lam_off = B.jrow, sd.row_h = B.row_h, sd.ground = B.parent < 0 ? 1 : 0;
            sd.rhs_off = B.jrow + (sd.role == PARENT ? 3 * M->hdr.nb : 0);
            if (B.type == LIN3) {
                sd.nl_slot = -1;
                const double sgn = sd.role == CHILD ? -1.0 : 1.0;  // Phi = c0 + a_p Q_p - a_c Q_c
                const double* a = B.p + (sd.role == CHILD ? 4 : 0);
                sd.lev[0] = sgn * (a[1] + a[2]);
                sd.lev[1] = sgn * a[0] - sd.lev[0] * c[0];
                sd.lev[2] = -sgn * a[2] - sd.lev[0] * c[1];
                sd.lev[3] = sgn * a[3] - sd.lev[0] * c[2];
            } else {
                // the row's 6-vector (n, f_dir); f_dir = sigma d vanishes identically for rows made of directions only
                // (sigma = a_rp + a_rd = 0: the rotation rows of hinge / universal joints), which then store n alone
                bool has_f = false;
                auto sigma_of = [](const double* a) { return a[1] + a[2]; };
                if (B.type == DOT) has_f = sigma_of(sd.role == CHILD ? B.p + 4 : B.p) != 0.0;
                else if (B.type == CLEN) has_f = true;
                else has_f = sd.role == CHILD ? sigma_of(B.p) != 0.0 : (sigma_of(B.p + 4) != 0.0 || sigma_of(B.p + 8) != 0.0);
                if (has_f) sd.ground |= 2;
                sd.nl_slot = nl;
                nl += has_f ? 6 : 3;
            }
            M->side.push_back(sd);
        }
        S.side_end = (int)M->side.size();
        if (nl > rnl) rnl = nl;
    }
    {  // nominal diagonal scale of every node of the joint-level system: a^T M4^-1 a summed over the two sides of a
        // row (state-dependent factors taken as 1); the kernel flags pivots that have cancelled against it
        const int nb = M->hdr.nb;
        std::vector<double> sc(nb, 0.0);
        auto quad = [&](int sgi, const double* a) {
            const double* W = Winv[sgi].data();
            double q = 0.0;
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 4; ++cc) {
                    const int lo = r < cc ? r : cc, hi = r < cc ? cc : r;
                    q += a[r] * W[lo * 4 - lo * (lo - 1) / 2 + (hi - lo)] * a[cc];
                }
            return std::fabs(q);
        };
        for (int b = 0; b < nblk; ++b) {
            const DevBlk& B = M->blk[b];
            double g = quad(B.child, B.p + 4);
            if (B.parent >= 0) g += quad(B.parent, B.p);
            const int node = B.jrow / 3;
            if (node < nb && g > sc[node]) sc[node] = g;
        }
        for (int I = 0; I < nb; ++I) {
            const float f = (float)sc[I];
            int bits;
            std::memcpy(&bits, &f, sizeof(bits));
            M->sym.push_back(bits);
        }
    }
    M->hdr.N = N, M->hdr.nblk = nblk, M->hdr.nside = (int)M->side.size();
    M->hdr.mj = mj, M->hdr.m = mj + 6 * N, M->hdr.spw = 32, M->hdr.rnl = rnl;  // spw is settled in build_blob
    if (d->nb_markers > 0) {
        if (d->marker_segment == nullptr || d->marker_position == nullptr) {
            delete M;
            return fail(BIONC_E_INVALID, "nb_markers > 0 without marker arrays");
        }
        std::vector<DevMarker> mk(d->nb_markers);
        for (int k = 0; k < d->nb_markers; ++k) {
            if (d->marker_segment[k] < 0 || d->marker_segment[k] >= N) {
                delete M;
                return fail(BIONC_E_INVALID, "marker attached to a segment outside [0, N)");
            }
            mk[k].segment = d->marker_segment[k], mk[k].pad = 0;
            for (int e = 0; e < 3; ++e) mk[k].pos[e] = d->marker_position[3 * k + e];
        }
        DeviceGuard guard(device);
        if (!guard.ok || cudaMalloc(&M->d_markers, mk.size() * sizeof(DevMarker)) != cudaSuccess ||
            cudaMemcpy(M->d_markers, mk.data(), mk.size() * sizeof(DevMarker), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            delete M;
            return fail(BIONC_E_CUDA, "marker table upload failed");
        }
        M->nb_markers = d->nb_markers;
    }
    layout_blob(M->hdr, M);
    *out = M;
    return BIONC_OK;
}

int bionc_cuda_model_destroy(BioncModel* M) {
    if (M == nullptr) return BIONC_OK;
    if (M->d_blob == nullptr && M->d_markers == nullptr && M->d_Q == nullptr) {  // the device was never touched
        delete M;
        return BIONC_OK;
    }
    DeviceGuard guard(M->device);
    for (auto& k : M->jit_kernels)
        if (k->mod != nullptr) jit::api().ModuleUnload(k->mod);
    cudaFree(M->d_blob);
    for (auto& j : M->jac) cudaFree(j.ptr), cudaFree(j.off), cudaFree(j.coef), cudaFree(j.c0);
    cudaFree(M->d_markers);
    cudaFree(M->d_Q), cudaFree(M->d_Qd), cudaFree(M->d_Qdd), cudaFree(M->d_lam), cudaFree(M->d_inertia), cudaFree(M->d_h);
    cudaFree(M->d_status), cudaFree(M->d_fvals);
    if (M->stream != nullptr) cudaStreamDestroy(M->stream);
    for (auto& ps : M->pipe)
        if (ps != nullptr) cudaStreamDestroy(ps);
    delete M;
    return BIONC_OK;
}

// shape of a call for the model-specialised kernels: what DynCall::prepare derives from the options (no upload)
static int spec_shape(const BioncModel* M, const BioncDynOptions* opt, int32_t systems_per_cta, jit::Variant& v) {
    DynOpts o;
    dyn_options(opt, o);
    const bool forces = opt != nullptr && opt->fext != nullptr && opt->fext->nb_forces > 0;
    const bool persys = opt != nullptr && opt->inertia != nullptr;  // only tested against nullptr
    FextArgs fa{};
    fa.tab = forces ? reinterpret_cast<const unsigned char*>(M) : nullptr;  // only tested against nullptr
    const int path = kernel_path(M, o, fa);
    int spc = systems_per_cta, ctas = 0;
    const int ftab = forces ? align16((int)(sizeof(DevFextTab) + (size_t)opt->fext->nb_forces * sizeof(DevFext))) : 0;
    if (spc == 0) {
        BIONC_ON_DEVICE(M);
        LaunchPlan plan = M->plan0;
        if (forces || persys) {
            const int rc = make_plan(M, ftab, persys, plan);
            if (rc != BIONC_OK) return rc;
        }
        spc = plan.spw, ctas = plan.ctas;
    } else if (spc == 16 || spc == 32) {  // no device: residency as on a B200 (228 KB of shared memory per SM)
        const DevHeader& h = M->hdr;
        const size_t bytes = (size_t)h.blob_bytes + ftab + (size_t)cta_smem_doubles(spc, h.N, h.nb, h.nslot, h.nslot2, h.rnl, persys) * sizeof(double);
        ctas = (int)(233472 / (bytes + 1024));
        if (ctas > 2048 / (32 * h.N)) ctas = 2048 / (32 * h.N);
        if (ctas > class_ctas(32 * h.N)) ctas = class_ctas(32 * h.N);
        if (ctas < 1) return fail(BIONC_E_TOO_LARGE, "model needs more shared memory per CTA than one SM has");
    }
    if (spc != 16 && spc != 32) return fail(BIONC_E_INVALID, "systems_per_cta must be 0 (the device's plan), 16 or 32");
    v = jit_variant(M, path, persys, forces, spc);
    v.ctas = ctas;
    {  // BIONC_SPEC_COMPACT=0/1 overrides the rule (experiment switch)
        const char* e = std::getenv("BIONC_SPEC_COMPACT");
        v.compact = e != nullptr ? std::atoi(e) != 0 : (M->hdr.serial_solve == 1 && M->hdr.mj > 0);
    }
    {  // Joint-level systems of more than four nodes keep the loops of their factorisation (measured: 4-link hinge chain,
        // 7 nodes, 1140 -> 956 ms per 1e6 x 100 steps; knee, 3 nodes, 151 -> 192 ms).  BIONC_SPEC_ROLL=0/1 overrides.
        const char* e = std::getenv("BIONC_SPEC_ROLL");
        v.roll_solve = e != nullptr ? std::atoi(e) != 0 : M->hdr.nb > 4;
        if (v.roll_solve || M->hdr.serial_solve != 1) v.compact = false;  // kept loops read the blob; parallel schedules re-read the pivot block
    }
    if (v.compact) {  // residency with the smaller working set (the class's CTA count still caps it)
        const DevHeader& h = M->hdr;
        const size_t bytes = (size_t)h.blob_bytes + ftab + (size_t)cta_smem_doubles(spc, h.N, h.nb, h.nslot, h.nslot2, h.rnl, persys) * sizeof(double) -
                             compact_savings(M, spc);
        const size_t sm_smem = M->sm_smem > 0 ? (size_t)M->sm_smem : (size_t)233472;
        int c = (int)(sm_smem / (bytes + 1024));
        if (c > 2048 / (32 * h.N)) c = 2048 / (32 * h.N);
        if (c > class_ctas(32 * h.N)) c = class_ctas(32 * h.N);
        if (c > v.ctas) v.ctas = c;
    }
    return BIONC_OK;
}

int bionc_cuda_model_spec_source(const BioncModel* M, const BioncDynOptions* opt, int32_t systems_per_cta, char* buf, size_t cap, size_t* needed) {
    if (M == nullptr) return fail(BIONC_E_INVALID, "null model");
    jit::Variant v;
    int rc = spec_shape(M, opt, systems_per_cta, v);
    if (rc != BIONC_OK) return rc;
    DevHeader h = M->hdr;
    h.spw = v.spc;
    const std::string src = jit::source(h, M->seg, M->blk, M->side, M->sym, v);
    if (needed != nullptr) *needed = src.size() + 1;
    if (buf != nullptr && cap > 0) {
        const size_t n = src.size() < cap - 1 ? src.size() : cap - 1;
        std::memcpy(buf, src.data(), n);
        buf[n] = '\0';
    }
    return BIONC_OK;
}

int bionc_cuda_model_specialize(BioncModel* M, const BioncDynOptions* opt, int32_t systems_per_cta, int32_t load) {
    if (M == nullptr) return fail(BIONC_E_INVALID, "null model");
    jit::Variant v;
    int rc = spec_shape(M, opt, load ? 0 : systems_per_cta, v);
    if (rc != BIONC_OK) return rc;
    {
        std::lock_guard<std::mutex> lock(M->jit_mu);
        for (auto& k : M->jit_kernels)
            if (k->v == v && (k->fn != nullptr || !load)) return BIONC_OK;
    }
    DevHeader h = M->hdr;
    h.spw = v.spc;
    std::unique_ptr<jit::Compiled> c(new jit::Compiled());
    std::string err = jit::compile(jit::source(h, M->seg, M->blk, M->side, M->sym, v), v, *c);
    if (!err.empty()) return fail(BIONC_E_CUDA, "model-specialised kernel: " + err);
    if (load) {
        BIONC_ON_DEVICE(M);
        CUDA_TRY(cudaFree(nullptr));  // the runtime's primary context is current on this thread from here on
        err = jit::load(*c, M->dev_smem_optin);
        if (!err.empty()) return fail(BIONC_E_CUDA, "model-specialised kernel: " + err);
    }
    std::lock_guard<std::mutex> lock(M->jit_mu);
    M->jit_kernels.push_back(std::move(c));
    return BIONC_OK;
}

int bionc_cuda_model_specialized(const BioncModel* Mc, int64_t* cubin_bytes) {
    if (Mc == nullptr) return 0;
    BioncModel* M = const_cast<BioncModel*>(Mc);
    std::lock_guard<std::mutex> lock(M->jit_mu);
    int n = 0;
    int64_t bytes = 0;
    for (auto& k : M->jit_kernels) n += k->fn != nullptr ? 1 : 0, bytes += (int64_t)k->cubin.size();
    if (cubin_bytes != nullptr) *cubin_bytes = bytes;
    return n;
}

int bionc_cuda_model_counts(const BioncModel* M, int32_t* nseg, int32_t* mj, int32_t* m) {
    if (M == nullptr) return fail(BIONC_E_INVALID, "null model");
    if (nseg) *nseg = M->hdr.N;
    if (mj) *mj = M->hdr.mj;
    if (m) *m = M->hdr.m;
    return BIONC_OK;
}

int bionc_cuda_model_joint_rows(const BioncModel* M, int32_t* jrow) {
    if (M == nullptr || jrow == nullptr) return fail(BIONC_E_INVALID, "null argument");
    for (size_t b = 0; b < M->blk.size(); ++b) jrow[b] = M->blk[b].jrow;
    return BIONC_OK;
}

int bionc_cuda_model_layout(const BioncModel* M, int32_t* out) {
    if (M == nullptr || out == nullptr) return fail(BIONC_E_INVALID, "null argument");
    const DevHeader& h = M->hdr;
    const int32_t v[11] = {h.N, h.mj, h.nb, h.nslot, h.nslot2, h.nlevel, h.rnl, h.nzero, M->plan0.spw, (int32_t)M->plan0.smem, h.serial_solve};
    std::memcpy(out, v, sizeof(v));
    return BIONC_OK;
}

// per-call launch state shared by the dynamics entry points
struct DynCall {
    DynOpts o{};
    FextCall fx;
    LaunchPlan plan;
    KernelSet ks{};
    int path = 0;
    const jit::Compiled* jit = nullptr;  // model-specialised integrator of this call shape, if the model has one
    const void* rk4_entry(const double* inertia) const {
        return ks.rk4[(inertia != nullptr ? 2 : 0) + (fx.args.tab != nullptr ? 1 : 0)];  // a force table implies path 2
    }
    int prepare(const BioncModel* M, const BioncDynOptions* opt, int64_t B, cudaStream_t st) {
        dyn_options(opt, o);
        int rc = fx.upload(M, opt != nullptr ? opt->fext : nullptr, st);
        if (rc != BIONC_OK) return rc;
        fx.args.step_stride = (long long)B * fx.args.nf * 6;
        plan = M->plan0;
        const bool persys = opt != nullptr && opt->inertia != nullptr;  // per-system constants take shared memory too
        if ((fx.args.tab != nullptr || persys) && (rc = make_plan(M, fx.args.tab_bytes, persys, plan)) != BIONC_OK) return rc;
        path = kernel_path(M, o, fx.args);
        ks = pick_kernels(path, 32 * M->hdr.N, plan.spw);
        jit = jit_lookup(M, jit_variant(M, path, persys, fx.args.tab != nullptr, plan.spw));
        return BIONC_OK;
    }
};

// The single-evaluation kernel is persistent: as many CTAs as fit the device at once, each walking tiles of systems, so
// the per-CTA set-up (model blob to shared memory) is paid once per CTA instead of once per 32 systems.
// grid of a persistent kernel: exactly the CTAs the device holds at once (a larger grid would run its surplus CTAs as a
// second wave after the first has walked all its tiles), or fewer if there are fewer tiles
static unsigned resident_grid(const BioncModel* M, const void* kernel, int threads, size_t smem, long long ntiles) {
    int per_sm = 0, sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->device);
    const long long resident = (long long)per_sm * sms;
    return (unsigned)(ntiles < resident ? ntiles : resident);
}
static unsigned fd_grid(const BioncModel* M, const DynCall& call, int64_t B) {
    return resident_grid(M, call.ks.fd, 32 * M->hdr.N, call.plan.smem, (B + call.plan.spw - 1) / call.plan.spw);
}

int bionc_cuda_forward_dynamics(const BioncModel* M, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                double* Qdd, double* lam, int32_t* status, int64_t B, void* stream) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || B < 0) return fail(BIONC_E_INVALID, "null argument");
    if (misaligned16(Q) || misaligned16(Qd) || misaligned16(Qdd)) return fail(BIONC_E_INVALID, "Q, Qdot and Qddot must be 16-byte aligned");
    if (B == 0) return BIONC_OK;
    BIONC_ON_DEVICE(M);
    cudaStream_t st = (cudaStream_t)stream;
    DynCall call;
    int rc = call.prepare(M, opt, B, st);
    if (rc != BIONC_OK) return rc;
    if (status != nullptr) CUDA_TRY(cudaMemsetAsync(status, 0, B * sizeof(int32_t), st));
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    const unsigned block = 32 * M->hdr.N;
    const unsigned char* blob = M->d_blob;
    int blob_bytes = M->blob_bytes;
    long long Bll = B;
    // lean path with everything 16-byte aligned: the kernel that moves whole system rows with the bulk-copy engine
    const void* rows_kernel = fd_rows_entry((int)block);
    // + the padding of the two row areas and a second buffer of Qdot rows
    const size_t rows_smem = call.plan.smem + (size_t)32 * (2 * 2 + 12 * M->hdr.N + 2) * sizeof(double);
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (kernel_path(M, call.o, call.fx.args) == 0 && call.plan.spw == 32 && inertia == nullptr && rows_kernel != nullptr &&
        aligned(Q) && aligned(Qd) && aligned(Qdd) && aligned(lam) && M->hdr.mj > 0 && M->hdr.m <= 9 * (M->hdr.nslot + M->hdr.nslot2) &&
        rows_smem <= (size_t)M->dev_smem_optin && (g_debug_poison.load() & 0x100) == 0) {
        if ((rc = ensure_smem(M->device, rows_kernel, M->dev_smem_optin)) != BIONC_OK) return rc;
        const unsigned grid = resident_grid(M, rows_kernel, block, rows_smem, (B + 31) / 32);
        void* args[] = {&blob, &blob_bytes, &Q, &Qd, &call.o, &Qdd, &lam, &status, &Bll};
        CUDA_TRY(cudaLaunchKernel(rows_kernel, dim3(grid), dim3(block), args, rows_smem, st));
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return BIONC_OK;
    }
    if ((rc = ensure_smem(M->device, call.ks.fd, M->dev_smem_optin)) != BIONC_OK) return rc;
    const unsigned grid = fd_grid(M, call, B);
    void* args[] = {&blob, &blob_bytes, &call.fx.args, &Q, &Qd, &inertia, &call.o, &Qdd, &lam, &status, &Bll};
    CUDA_TRY(cudaLaunchKernel(call.ks.fd, dim3(grid), dim3(block), args, call.plan.smem, st));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

size_t bionc_cuda_dense_workspace_bytes(const BioncModel* M, int64_t count) {
    if (M == nullptr || count < 1) return 0;
    const size_t Mdim = (size_t)12 * M->hdr.N + (size_t)(M->hdr.mj + 6 * M->hdr.N);
    const size_t per = (Mdim * (Mdim + 1) + Mdim) * sizeof(double);
    const int64_t slots = count < 148 * 8 ? count : 148 * 8;
    return per * (size_t)slots;
}

int bionc_cuda_forward_dynamics_dense(const BioncModel* M, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                      const int64_t* select, int64_t count, double* Qdd, double* lam, int32_t* status,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || count < 0) return fail(BIONC_E_INVALID, "null argument");
    if (count == 0) return BIONC_OK;
    const size_t per = bionc_cuda_dense_workspace_bytes(M, 1);
    if (workspace == nullptr || workspace_bytes < per) return fail(BIONC_E_INVALID, "workspace smaller than bionc_cuda_dense_workspace_bytes(model, 1)");
    BIONC_ON_DEVICE(M);
    cudaStream_t st = (cudaStream_t)stream;
    DynOpts o;
    dyn_options(opt, o);
    FextCall fx;
    int rc = fx.upload(M, opt != nullptr ? opt->fext : nullptr, st);
    if (rc != BIONC_OK) return rc;
    long long slots = (long long)(workspace_bytes / per);
    if (slots > count) slots = count;
    if (slots > 148 * 8) slots = 148 * 8;
    EvalArgs ea{0, M->hdr.N, M->hdr.nblk, M->hdr.mj, M->hdr.m};
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    dense_kkt_kernel<<<(unsigned)slots, kDenseThreads, 0, st>>>(
        M->d_blob, fx.args, ea, Q, Qd, inertia, o, reinterpret_cast<const long long*>(select), (long long)count, Qdd, lam, status,
        static_cast<double*>(workspace));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

// one launch of the fused integrator: steps step0 .. step0 + n_steps - 1 of a run (h_steps, traj and the force slices
// are indexed by the global step)
static int launch_rk4(const BioncModel* M, DynCall& call, double* Q, double* Qd, double h, const double* h_steps, long long step0,
                      long long n_steps, int normalize, const double* inertia, double* traj, long long traj_every, int32_t* status,
                      long long B, cudaStream_t st) {
    const unsigned grid = (unsigned)((B + call.plan.spw - 1) / call.plan.spw), block = 32 * M->hdr.N;
    const unsigned char* blob = M->d_blob;
    int blob_bytes = M->blob_bytes;
    void* args[] = {&blob, &blob_bytes, &call.fx.args, &Q, &Qd, &inertia, &call.o, &h, &h_steps, &n_steps, &step0, &normalize, &traj, &traj_every, &status, &B};
    if (call.jit != nullptr) {  // this model's own build of the same kernel (bionc_cuda_model_specialize)
        const size_t smem = call.plan.smem - (call.jit->v.compact ? compact_savings(M, call.plan.spw) : 0);
        const CUresult r = jit::api().LaunchKernel(call.jit->fn, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)st, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(BIONC_E_CUDA, "cuLaunchKernel (model-specialised integrator): " + jit::cu_error(r));
        g_launches++;
        return BIONC_OK;
    }
    CUDA_TRY(cudaLaunchKernel(call.rk4_entry(inertia), dim3(grid), dim3(block), args, call.plan.smem, st));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

int bionc_cuda_rk4(const BioncModel* M, double* Q, double* Qd, double h, const double* h_steps, int64_t n_steps,
                   int32_t normalize, const BioncDynOptions* opt, double* traj, int64_t traj_every, double* lam_last,
                   int32_t* status, int64_t B, void* stream) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || B < 0 || n_steps < 0) return fail(BIONC_E_INVALID, "null argument");
    if (traj != nullptr && traj_every < 1) return fail(BIONC_E_INVALID, "traj_every must be >= 1");
    if (misaligned16(Q) || misaligned16(Qd) || misaligned16(traj)) return fail(BIONC_E_INVALID, "Q, Qdot and traj must be 16-byte aligned");
    if (B == 0 || n_steps == 0) return BIONC_OK;
    BIONC_ON_DEVICE(M);
    cudaStream_t st = (cudaStream_t)stream;
    DynCall call;
    int rc = call.prepare(M, opt, B, st);
    if (rc != BIONC_OK) return rc;
    if (status != nullptr) CUDA_TRY(cudaMemsetAsync(status, 0, B * sizeof(int32_t), st));
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    if (traj == nullptr) traj_every = 1;
    if (call.jit == nullptr && (rc = ensure_smem(M->device, call.rk4_entry(inertia), M->dev_smem_optin)) != BIONC_OK) return rc;
    if (lam_last == nullptr) return launch_rk4(M, call, Q, Qd, h, h_steps, 0, n_steps, normalize, inertia, traj, traj_every, status, B, st);
    // The multipliers of the first stage of the last step are those of forward_dynamics at the state before that step:
    // n_steps - 1 fused steps, one forward_dynamics launch for lambda (its accelerations go to a scratch array), the last
    // step.  The integrator kernel itself carries no multiplier output path.
    if (n_steps > 1 && (rc = launch_rk4(M, call, Q, Qd, h, h_steps, 0, n_steps - 1, normalize, inertia, traj, traj_every, status, B, st)) != BIONC_OK) return rc;
    {
        if ((rc = ensure_smem(M->device, call.ks.fd, M->dev_smem_optin)) != BIONC_OK) return rc;
        double* scratch = nullptr;
        CUDA_TRY(cudaMallocAsync(&scratch, (size_t)B * 12 * M->hdr.N * sizeof(double), st));
        FextArgs fxl = call.fx.args;  // the force slice of the last step
        if (fxl.vals != nullptr) fxl.vals += (n_steps - 1 < fxl.steps ? n_steps - 1 : fxl.steps - 1) * fxl.step_stride;
        const unsigned grid = fd_grid(M, call, B), block = 32 * M->hdr.N;
        const unsigned char* blob = M->d_blob;
        int blob_bytes = M->blob_bytes;
        const double *Qc = Q, *Qdc = Qd;
        int32_t* no_status = nullptr;  // flags of this evaluation are raised by the last step as well
        long long Bll = B;
        void* args[] = {&blob, &blob_bytes, &fxl, &Qc, &Qdc, &inertia, &call.o, &scratch, &lam_last, &no_status, &Bll};
        const cudaError_t e = cudaLaunchKernel(call.ks.fd, dim3(grid), dim3(block), args, call.plan.smem, st);
        cudaFreeAsync(scratch, st);
        if (e != cudaSuccess) return fail(BIONC_E_CUDA, std::string("cudaLaunchKernel: ") + cudaGetErrorString(e));
        g_launches++;
    }
    return launch_rk4(M, call, Q, Qd, h, h_steps, n_steps - 1, 1, normalize, inertia, traj, traj_every, status, B, st);
}

int bionc_cuda_eval(const BioncModel* M, int32_t which, const double* in, const BioncDynOptions* opt, double* out,
                    int64_t B, void* stream) {
    if (M == nullptr || in == nullptr || out == nullptr || B < 0 || which < 0 || which > 12) return fail(BIONC_E_INVALID, "bad argument");
    if (misaligned16(in) || misaligned16(out)) return fail(BIONC_E_INVALID, "in and out must be 16-byte aligned");
    if (B == 0) return BIONC_OK;
    BIONC_ON_DEVICE(M);
    const int N = M->hdr.N, n = 12 * N, mj = M->hdr.mj, m = M->hdr.m;
    cudaStream_t st = (cudaStream_t)stream;
    if (which >= BIONC_EVAL_COM) {  // post-processing evaluators: every output element is written, no zero fill
        if (which == BIONC_EVAL_COM || which == BIONC_EVAL_MARKERS) {
            const int nitem = which == BIONC_EVAL_COM ? N : M->nb_markers;
            if (nitem == 0) return BIONC_OK;
            const long long total = B * (long long)nitem;
            long long grid = (total + 255) / 256;
            if (grid > 148LL * 32) grid = 148LL * 32;
            point_kernel<<<(unsigned)grid, 256, 0, st>>>(M->d_blob, M->d_markers, which, N, nitem, in, out, B);
        } else {
            int G = 1;  // lanes per system: segments rounded up to a power of two
            while (G < N) G <<= 1;
            const long long threads = B * (long long)G;
            long long grid = (threads + 255) / 256;
            if (grid > 148LL * 32) grid = 148LL * 32;
            if (which == BIONC_EVAL_KINETIC) reduce_kernel<10><<<(unsigned)grid, 256, 0, st>>>(M->d_blob, N, M->hdr.nblk, G, in, out, B);
            else if (which == BIONC_EVAL_POTENTIAL) reduce_kernel<11><<<(unsigned)grid, 256, 0, st>>>(M->d_blob, N, M->hdr.nblk, G, in, out, B);
            else reduce_kernel<12><<<(unsigned)grid, 256, 0, st>>>(M->d_blob, N, M->hdr.nblk, G, in, out, B);
        }
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return BIONC_OK;
    }
    int rows = 0, cols = 1;
    switch (which) {
        case BIONC_EVAL_PHI_R: rows = 6 * N; break;
        case BIONC_EVAL_K_R: rows = 6 * N, cols = n; break;
        case BIONC_EVAL_PHI_J: rows = mj; break;
        case BIONC_EVAL_K_J: rows = mj, cols = n; break;
        case BIONC_EVAL_PHI_H: case BIONC_EVAL_BIAS_H: rows = m; break;
        case BIONC_EVAL_K_H: rows = m, cols = n; break;
        case BIONC_EVAL_FEXT: rows = n; break;
    }
    const int out_per_sys = rows * cols;
    if (out_per_sys == 0) return BIONC_OK;
    FextCall fx;
    if (which == BIONC_EVAL_FEXT) {
        int rc = fx.upload(M, opt != nullptr ? opt->fext : nullptr, st);
        if (rc != BIONC_OK) return rc;
        fx.args.step_stride = (long long)B * fx.args.nf * 6;
    }
    if (which == BIONC_EVAL_FEXT) {
        // natural external forces: tile of T systems (T even: every full tile is a 16-byte multiple for the bulk copies)
        EvalArgs ea{which, N, M->hdr.nblk, mj, m};
        long long T = 65536 / (16LL * (n + rows));
        T = T > 64 ? 64 : (T < 2 ? 2 : T & ~1LL);
        if (T > B) T = B;
        const size_t smem = 128 + 2 * (size_t)T * (n + rows) * sizeof(double);
        int rc = ensure_smem(M->device, (const void*)eval_vec_kernel, M->dev_smem_optin);
        if (rc != BIONC_OK) return rc;
        const unsigned grid = resident_grid(M, (const void*)eval_vec_kernel, kEvalThreads, smem, (B + T - 1) / T);
        eval_vec_kernel<<<grid, kEvalThreads, smem, st>>>(M->d_blob, fx.args, ea, in, out, (long long)B, (int)T, rows);
    } else if (cols == 1) {
        // constraint vectors: 32 systems per tile, lane = system, one item per instruction stream
        EvalArgs ea{which, N, M->hdr.nblk, mj, m};
        const size_t smem = ((size_t)kItemSystems * (n + 1) + (size_t)kItemSystems * (rows | 1)) * sizeof(double);
        if (smem > (size_t)M->dev_smem_optin) return fail(BIONC_E_TOO_LARGE, "model too large for the constraint evaluator tile");
        int rc = ensure_smem(M->device, (const void*)items_kernel, M->dev_smem_optin);
        if (rc != BIONC_OK) return rc;
        const unsigned grid = resident_grid(M, (const void*)items_kernel, kItemThreads, smem, (B + kItemSystems - 1) / kItemSystems);
        items_kernel<<<grid, kItemThreads, smem, st>>>(M->d_blob, ea, rows, in, out, (long long)B);
    } else {
        const BioncModel::Jac& J = M->jac[which == BIONC_EVAL_K_R ? 0 : (which == BIONC_EVAL_K_J ? 1 : 2)];
        const size_t smem = (size_t)kJacSystems * n * sizeof(double);
        const unsigned grid = resident_grid(M, (const void*)jacobian_kernel, kJacThreads, smem, (B + kJacSystems - 1) / kJacSystems);
        jacobian_kernel<<<grid, kJacThreads, smem, st>>>(J.ptr, J.off, J.coef, J.c0, J.E, n, in, out, (long long)B);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

int bionc_cuda_mass_matrix(const BioncModel* M, double* G) {
    if (M == nullptr || G == nullptr) return fail(BIONC_E_INVALID, "null argument");
    const int N = M->hdr.N, n = 12 * N;
    std::memset(G, 0, sizeof(double) * n * n);
    for (int i = 0; i < N; ++i) {
        const double* p = M->seg[i].mcj;
        const double m = p[0], c0 = p[1], c1 = p[2], c2 = p[3], J00 = p[4], J01 = p[5], J02 = p[6], J11 = p[7], J12 = p[8], J22 = p[9];
        const double M4[4][4] = {{J00, m * c0 + J01, -J01, J02},
                                 {m * c0 + J01, m + 2 * m * c1 + J11, -(m * c1 + J11), m * c2 + J12},
                                 {-J01, -(m * c1 + J11), J11, -J12},
                                 {J02, m * c2 + J12, -J12, J22}};
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                for (int k = 0; k < 3; ++k) G[(size_t)(12 * i + 3 * a + k) * n + 12 * i + 3 * b + k] = M4[a][b];
    }
    return BIONC_OK;
}

// ------------------------------------------------------------------------------------------ host path
}  // extern "C"
// (re)allocate one scratch array; on failure the pointer stays null and the capacity 0
template <typename T>
static int regrow(T*& p, size_t count) {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    return BIONC_OK;
}
extern "C" {

// host_mu must be held
static int ensure_host_scratch(BioncModel* M, long long B, long long nh, bool want_inertia, long long fvals_doubles) {
    if (M->stream == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking));
    for (auto& ps : M->pipe)
        if (ps == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    const size_t n = 12 * (size_t)M->hdr.N;
    int rc;
    if (B > M->cap_B) {
        M->cap_B = 0;  // stays 0 unless every allocation below succeeds
        if (M->d_inertia != nullptr) cudaFree(M->d_inertia), M->d_inertia = nullptr;
        if ((rc = regrow(M->d_Q, B * n)) != BIONC_OK || (rc = regrow(M->d_Qd, B * n)) != BIONC_OK ||
            (rc = regrow(M->d_Qdd, B * n)) != BIONC_OK || (rc = regrow(M->d_lam, B * (size_t)M->hdr.m)) != BIONC_OK ||
            (rc = regrow(M->d_status, (size_t)B)) != BIONC_OK)
            return rc;
        M->cap_B = B;
    }
    if (want_inertia && M->d_inertia == nullptr && (rc = regrow(M->d_inertia, M->cap_B * (size_t)M->hdr.N * 10)) != BIONC_OK) return rc;
    if (nh > M->cap_h) {
        M->cap_h = 0;
        if ((rc = regrow(M->d_h, (size_t)nh)) != BIONC_OK) return rc;
        M->cap_h = nh;
    }
    if (fvals_doubles > M->cap_fvals) {
        M->cap_fvals = 0;
        if ((rc = regrow(M->d_fvals, (size_t)fvals_doubles)) != BIONC_OK) return rc;
        M->cap_fvals = fvals_doubles;
    }
    return BIONC_OK;
}

// Chunking of the host-buffer entry points: the batch is cut into up to 8 chunks that go round-robin
// through three streams, so the H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the kernel
// of chunk i (PCIe is full duplex); chunks are multiples of 4096 systems.
static int64_t host_chunk(int64_t B) {
    const int64_t min_chunk = 65536;
    int64_t n = (B + min_chunk - 1) / min_chunk;
    if (n > 8) n = 8;
    if (n < 1) n = 1;
    int64_t c = (B + n - 1) / n;
    return (c + 4095) / 4096 * 4096;
}

// drain the chunk pipeline (also on the error paths: no copy into the caller's buffers may be left in flight)
static int drain(BioncModel* M, int rc) {
    for (auto& ps : M->pipe)
        if (ps != nullptr) {
            const cudaError_t e = cudaStreamSynchronize(ps);
            if (e != cudaSuccess && rc == BIONC_OK) rc = fail(BIONC_E_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
        }
    return rc;
}
#define PIPE_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return drain(M, fail(BIONC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__))); \
    } while (0)

// per-chunk options of the host path: device copies of the per-system inertia / force values of the chunk
struct HostChunkOpts {
    BioncDynOptions o{};
    BioncFextDesc fx{};
};
static int stage_chunk_options(BioncModel* M, const BioncDynOptions* opt, int64_t B, int64_t lo, int64_t b, cudaStream_t st, HostChunkOpts& out) {
    out.o = opt != nullptr ? *opt : BioncDynOptions{};
    if (opt != nullptr && opt->inertia != nullptr) {
        const size_t ni = (size_t)M->hdr.N * 10;
        CUDA_TRY(cudaMemcpyAsync(M->d_inertia + lo * ni, opt->inertia + lo * ni, (size_t)b * ni * sizeof(double), cudaMemcpyHostToDevice, st));
        out.o.inertia = M->d_inertia + lo * ni;
    }
    if (opt != nullptr && opt->fext != nullptr && opt->fext->values != nullptr && opt->fext->nb_forces > 0) {
        // (steps, B, F, 6) on the host -> (steps, b, F, 6) of this chunk, slices packed back to back in the scratch
        out.fx = *opt->fext;
        const size_t per = (size_t)opt->fext->nb_forces * 6;
        const int64_t steps = opt->fext->values_steps > 1 ? opt->fext->values_steps : 1;
        double* dst = M->d_fvals + (size_t)lo * per * steps;
        for (int64_t k = 0; k < steps; ++k)
            CUDA_TRY(cudaMemcpyAsync(dst + (size_t)k * b * per, opt->fext->values + ((size_t)k * B + lo) * per, (size_t)b * per * sizeof(double),
                                     cudaMemcpyHostToDevice, st));
        out.fx.values = dst;
        out.o.fext = &out.fx;
    }
    return BIONC_OK;
}
static long long host_fvals_doubles(const BioncDynOptions* opt, int64_t B) {
    if (opt == nullptr || opt->fext == nullptr || opt->fext->values == nullptr || opt->fext->nb_forces <= 0) return 0;
    const int64_t steps = opt->fext->values_steps > 1 ? opt->fext->values_steps : 1;
    return (long long)B * opt->fext->nb_forces * 6 * steps;
}

int bionc_cuda_forward_dynamics_host(BioncModel* M, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                     double* Qdd, double* lam, int32_t* status, int64_t B) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || B < 0) return fail(BIONC_E_INVALID, "null argument");
    if (B == 0) return BIONC_OK;
    BIONC_ON_DEVICE(M);
    std::lock_guard<std::mutex> lock(M->host_mu);
    const bool per_sys = opt != nullptr && opt->inertia != nullptr;
    int rc = ensure_host_scratch(M, B, 0, per_sys, host_fvals_doubles(opt, B));
    if (rc != BIONC_OK) return rc;
    const size_t n = 12 * (size_t)M->hdr.N, m = (size_t)M->hdr.m;
    const int64_t chunk = host_chunk(B);
    int k = 0;
    for (int64_t lo = 0; lo < B; lo += chunk, ++k) {
        const int64_t b = (B - lo < chunk) ? B - lo : chunk;
        cudaStream_t st = M->pipe[k % 3];
        const size_t nb = (size_t)b * n * sizeof(double);
        PIPE_TRY(cudaMemcpyAsync(M->d_Q + lo * n, Q + lo * n, nb, cudaMemcpyHostToDevice, st));
        PIPE_TRY(cudaMemcpyAsync(M->d_Qd + lo * n, Qd + lo * n, nb, cudaMemcpyHostToDevice, st));
        HostChunkOpts co;
        if ((rc = stage_chunk_options(M, opt, B, lo, b, st, co)) != BIONC_OK) return drain(M, rc);
        rc = bionc_cuda_forward_dynamics(M, M->d_Q + lo * n, M->d_Qd + lo * n, &co.o, M->d_Qdd + lo * n,
                                         lam != nullptr ? M->d_lam + lo * m : nullptr, status != nullptr ? M->d_status + lo : nullptr, b, st);
        if (rc != BIONC_OK) return drain(M, rc);
        PIPE_TRY(cudaMemcpyAsync(Qdd + lo * n, M->d_Qdd + lo * n, nb, cudaMemcpyDeviceToHost, st));
        if (lam != nullptr) PIPE_TRY(cudaMemcpyAsync(lam + lo * m, M->d_lam + lo * m, (size_t)b * m * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (status != nullptr) PIPE_TRY(cudaMemcpyAsync(status + lo, M->d_status + lo, (size_t)b * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    return drain(M, BIONC_OK);
}

int bionc_cuda_rk4_host(BioncModel* M, double* Q, double* Qd, double h, const double* h_steps, int64_t n_steps,
                        int32_t normalize, const BioncDynOptions* opt, int32_t* status, int64_t B) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || B < 0 || n_steps < 0) return fail(BIONC_E_INVALID, "null argument");
    if (B == 0 || n_steps == 0) return BIONC_OK;
    BIONC_ON_DEVICE(M);
    std::lock_guard<std::mutex> lock(M->host_mu);
    const bool per_sys = opt != nullptr && opt->inertia != nullptr;
    int rc = ensure_host_scratch(M, B, h_steps != nullptr ? n_steps : 0, per_sys, host_fvals_doubles(opt, B));
    if (rc != BIONC_OK) return rc;
    const size_t n = 12 * (size_t)M->hdr.N;
    if (h_steps != nullptr) {
        CUDA_TRY(cudaMemcpyAsync(M->d_h, h_steps, n_steps * sizeof(double), cudaMemcpyHostToDevice, M->stream));
        CUDA_TRY(cudaStreamSynchronize(M->stream));
    }
    const int64_t chunk = host_chunk(B);
    int k = 0;
    for (int64_t lo = 0; lo < B; lo += chunk, ++k) {
        const int64_t b = (B - lo < chunk) ? B - lo : chunk;
        cudaStream_t st = M->pipe[k % 3];
        const size_t nb = (size_t)b * n * sizeof(double);
        PIPE_TRY(cudaMemcpyAsync(M->d_Q + lo * n, Q + lo * n, nb, cudaMemcpyHostToDevice, st));
        PIPE_TRY(cudaMemcpyAsync(M->d_Qd + lo * n, Qd + lo * n, nb, cudaMemcpyHostToDevice, st));
        HostChunkOpts co;
        if ((rc = stage_chunk_options(M, opt, B, lo, b, st, co)) != BIONC_OK) return drain(M, rc);
        rc = bionc_cuda_rk4(M, M->d_Q + lo * n, M->d_Qd + lo * n, h, h_steps != nullptr ? M->d_h : nullptr, n_steps, normalize,
                            &co.o, nullptr, 1, nullptr, status != nullptr ? M->d_status + lo : nullptr, b, st);
        if (rc != BIONC_OK) return drain(M, rc);
        PIPE_TRY(cudaMemcpyAsync(Q + lo * n, M->d_Q + lo * n, nb, cudaMemcpyDeviceToHost, st));
        PIPE_TRY(cudaMemcpyAsync(Qd + lo * n, M->d_Qd + lo * n, nb, cudaMemcpyDeviceToHost, st));
        if (status != nullptr) PIPE_TRY(cudaMemcpyAsync(status + lo, M->d_status + lo, (size_t)b * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    return drain(M, BIONC_OK);
}

// FP64 roofline denominator: a DFMA-only kernel (8 independent accumulator chains per thread,
// 2048 threads per SM resident) timed with CUDA events; MEASURED_PEAKS.json has no FP64 figure.
int bionc_cuda_measure_fp64_peak(int device, int repeats, double* tflops_best, double* tflops_sustained) {
    if (tflops_best == nullptr) return fail(BIONC_E_INVALID, "null argument");
    DeviceGuard guard__(device);
    if (!guard__.ok) return fail(BIONC_E_CUDA, "cudaSetDevice failed");
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int block = 256, grid = sms * 8, iters = 1 << 14;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, (size_t)grid * block * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const double flops = 2.0 * 8.0 * (double)iters * (double)grid * block;
    double best = 0.0, total_ms = 0.0;
    if (repeats < 3) repeats = 3;
    for (int r = 0; r < repeats + 2; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, block>>>(d_out, iters, 1.0000001, 1e-9);
        g_launches++;
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) {  // two warm-up launches
            const double tf = flops / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
            total_ms += ms;
        }
    }
    *tflops_best = best;
    if (tflops_sustained != nullptr) *tflops_sustained = flops * repeats / (total_ms * 1e-3) / 1e12;
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    CUDA_TRY(cudaFree(d_out));
    return BIONC_OK;
}

int bionc_cuda_debug_poison(int32_t mask) {
    g_debug_poison.store(mask);
    return BIONC_OK;
}
int64_t bionc_cuda_launch_count(void) { return g_launches.load(); }
const char* bionc_cuda_last_error(void) { return g_last_error.c_str(); }
const char* bionc_cuda_version(void) { return "bionc_cuda 0.1.0 (sm_100a, fp64)"; }

}  // extern "C"
