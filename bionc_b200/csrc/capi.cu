// C ABI (include/bionc_cuda.h): model lowering to the device blob, launch logic, host-buffer path.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bionc_cuda.h"
#include "kernels.cuh"

using namespace bionc;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

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

struct BioncModel {
    int device = 0;
    DevHeader hdr{};
    std::vector<DevSeg> seg;
    std::vector<DevBlk> blk;
    std::vector<DevSide> side;
    std::vector<int> sym;  // symbolic factorisation tables (see DevHeader::off_sym)
    int nb_markers = 0;
    DevMarker* d_markers = nullptr;  // marker table (device), outside the blob: the dynamics kernels never see it
    // device blob for the current external-force set (rebuilt when the set changes)
    std::vector<unsigned char> fext_key;
    unsigned char* d_blob = nullptr;
    int blob_bytes = 0;
    bool blob_valid = false;
    size_t smem_bytes = 0;
    std::mutex mu;
    // host-path scratch
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // chunk pipeline of the host-buffer entry points
    double *d_Q = nullptr, *d_Qd = nullptr, *d_Qdd = nullptr, *d_lam = nullptr, *d_inertia = nullptr, *d_h = nullptr;
    int* d_status = nullptr;
    long long cap_B = 0, cap_h = 0;
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
    // padding rows sit at the end of the node that received the last single rows
    for (int I = 0; I < nb; ++I)
        if (mixed[I]) {
            int rows = (int)node_blocks[I].size();
            for (int r = rows; r < 3; ++r) unit.push_back(9 * slot_tab[newidx[I] * nb + newidx[I]] + 4 * r);
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
    M->hdr.nzero = (int)zero.size(), M->hdr.ncomb = (int)comb.size() / 2, M->hdr.nunit = (int)unit.size();
    M->hdr.ncol = (int)colrow.size();
    return mj;
}


int build_blob(BioncModel* M, const BioncFextDesc* fx) {
    // key = raw bytes of the force description
    std::vector<unsigned char> key;
    std::vector<DevFext> fext;
    const int N = M->hdr.N;
    if (fx != nullptr && fx->nb_forces > 0) {
        for (int i = 0; i < fx->nb_forces; ++i) {
            if (fx->segment[i] < 0 || fx->segment[i] >= N) return fail(BIONC_E_INVALID, "external force on a segment that does not exist");
            if (fx->kind[i] < 0 || fx->kind[i] > 3) return fail(BIONC_E_INVALID, "unknown external force kind");
        }
        key.resize(sizeof(int) + fx->nb_forces * (2 * sizeof(int) + kFextNParam * sizeof(double)));
        unsigned char* k = key.data();
        std::memcpy(k, &fx->nb_forces, sizeof(int)), k += sizeof(int);
        std::memcpy(k, fx->segment, fx->nb_forces * sizeof(int)), k += fx->nb_forces * sizeof(int);
        std::memcpy(k, fx->kind, fx->nb_forces * sizeof(int)), k += fx->nb_forces * sizeof(int);
        std::memcpy(k, fx->param, fx->nb_forces * kFextNParam * sizeof(double));
    }
    if (M->blob_valid && key == M->fext_key) return BIONC_OK;

    std::vector<DevSeg> seg = M->seg;
    for (int s = 0; s < N; ++s) {  // forces grouped by segment, insertion order kept inside a segment
        seg[s].fext_begin = (int)fext.size();
        if (fx != nullptr)
            for (int i = 0; i < fx->nb_forces; ++i)
                if (fx->segment[i] == s) {
                    DevFext f{};
                    std::memcpy(f.p, fx->param + (size_t)i * kFextNParam, kFextNParam * sizeof(double));
                    f.segment = s, f.kind = fx->kind[i];
                    fext.push_back(f);
                }
        seg[s].fext_end = (int)fext.size();
    }
    DevHeader h = M->hdr;
    h.nfext = (int)fext.size();
    h.off_seg = align16(sizeof(DevHeader));
    h.off_blk = align16(h.off_seg + (int)(seg.size() * sizeof(DevSeg)));
    h.off_side = align16(h.off_blk + (int)(M->blk.size() * sizeof(DevBlk)));
    h.off_fext = align16(h.off_side + (int)(M->side.size() * sizeof(DevSide)));
    h.off_sym = align16(h.off_fext + (int)(fext.size() * sizeof(DevFext)));
    h.blob_bytes = align16(h.off_sym + (int)(M->sym.size() * sizeof(int)));
    std::vector<unsigned char> blob(h.blob_bytes, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    std::memcpy(blob.data() + h.off_seg, seg.data(), seg.size() * sizeof(DevSeg));
    if (!M->blk.empty()) std::memcpy(blob.data() + h.off_blk, M->blk.data(), M->blk.size() * sizeof(DevBlk));
    if (!M->side.empty()) std::memcpy(blob.data() + h.off_side, M->side.data(), M->side.size() * sizeof(DevSide));
    if (!fext.empty()) std::memcpy(blob.data() + h.off_fext, fext.data(), fext.size() * sizeof(DevFext));
    if (!M->sym.empty()) std::memcpy(blob.data() + h.off_sym, M->sym.data(), M->sym.size() * sizeof(int));

    // shared-memory budget of one CTA (32 systems, one warp per segment)
    int dev_smem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, M->device));
    // Systems per CTA: 32 (every lane busy) or 16 (upper half-warps idle).  Pick what keeps more systems
    // resident per SM: large per-system working sets (many joint rows) can fit three 16-system CTAs where
    // only one 32-system CTA fits.
    int sm_smem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, M->device));
    size_t cta_bytes = 0;
    int best_spc = 0, best_resident = 0;
    for (int spc = 32; spc >= 16; spc >>= 1) {
        const size_t bytes = h.blob_bytes + (size_t)cta_smem_doubles(spc, h.N, h.nb, h.nslot, h.nslot2, h.rnl) * sizeof(double);
        if (bytes > (size_t)dev_smem) continue;
        int ctas = (int)((size_t)sm_smem / (bytes + 1024));  // 1 KB per CTA is reserved by the driver
        const int by_threads = 2048 / (32 * h.N);
        if (ctas > by_threads) ctas = by_threads;
        if (ctas > 32) ctas = 32;
        if (ctas * spc > best_resident) best_resident = ctas * spc, best_spc = spc, cta_bytes = bytes;
    }
    if (best_spc == 0) return fail(BIONC_E_TOO_LARGE, "model needs more shared memory per CTA than one SM has");
    h.spw = best_spc;

    CUDA_TRY(cudaSetDevice(M->device));
    if (M->d_blob != nullptr && M->blob_bytes < h.blob_bytes) {
        CUDA_TRY(cudaFree(M->d_blob));
        M->d_blob = nullptr;
    }
    if (M->d_blob == nullptr) CUDA_TRY(cudaMalloc(&M->d_blob, h.blob_bytes));
    // synchronous copy: the blob is a few KB and changes only when the force set changes
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(M->d_blob, blob.data(), h.blob_bytes, cudaMemcpyHostToDevice));
    M->blob_bytes = h.blob_bytes;
    M->hdr = h;
    M->smem_bytes = cta_bytes;
    M->fext_key.swap(key);
    M->blob_valid = true;
    return BIONC_OK;
}

// kernel instantiations by CTA-size class (threads = 32 x segments) and path (lean / general)
struct KernelSet {
    const void* fd;
    const void* rk4;
};
#define BIONC_PICK(T, S)                                                                                          \
    (general ? KernelSet{(const void*)forward_dynamics_kernel<true, T, S>, (const void*)rk4_kernel<true, T, S>}    \
             : KernelSet{(const void*)forward_dynamics_kernel<false, T, S>, (const void*)rk4_kernel<false, T, S>})
KernelSet pick_kernels(bool general, int threads, int spc) {
    if (spc == 32) {
        if (threads <= 32) return BIONC_PICK(32, 32);
        if (threads <= 64) return BIONC_PICK(64, 32);
        if (threads <= 96) return BIONC_PICK(96, 32);
        if (threads <= 128) return BIONC_PICK(128, 32);
        if (threads <= 256) return BIONC_PICK(256, 32);
        if (threads <= 384) return BIONC_PICK(384, 32);
        if (threads <= 512) return BIONC_PICK(512, 32);
        return BIONC_PICK(1024, 32);
    }
    // 16 systems per CTA (upper half-warps mirror the lower ones): models whose working set for 32 systems
    // would leave fewer systems resident per SM
    if (threads <= 128) return BIONC_PICK(128, 16);
    if (threads <= 256) return BIONC_PICK(256, 16);
    return BIONC_PICK(1024, 16);
}

int ensure_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, size_t>> done;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& d : done)
        if (d.first == kernel && d.second >= bytes) return BIONC_OK;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // as much shared memory as the SM has: residency is limited by the per-CTA working sets
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    done.emplace_back(kernel, bytes);
    return BIONC_OK;
}

int prepare(BioncModel* M, const BioncDynOptions* opt, DynOpts& o) {
    o.use_stab = (opt != nullptr && opt->use_stabilization) ? 1 : 0;
    o.alpha = opt != nullptr ? opt->alpha : 0.0;
    o.beta = opt != nullptr ? opt->beta : 0.0;
    return build_blob(M, opt != nullptr ? opt->fext : nullptr);
}

// the lean kernel instantiation covers models made of point-to-point (LIN3) blocks only, without
// external forces or Baumgarte stabilisation; everything else takes the general one
bool needs_general(const BioncModel* M, const DynOpts& o) { return M->hdr.rnl > 0 || M->hdr.nfext > 0 || o.use_stab != 0; }

unsigned grid_for(const BioncModel* M, long long B) { return (unsigned)((B + M->hdr.spw - 1) / M->hdr.spw); }

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
        double W[4][4];
        if (!invert4(M4, W)) {
            delete M;
            return fail(BIONC_E_INVALID, "singular segment mass matrix");  // reference: LinAlgError at solve time
        }
        int k = 0;
        for (int r = 0; r < 4; ++r)
            for (int cc = r; cc < 4; ++cc) S.W[k++] = 0.5 * (W[r][cc] + W[cc][r]);
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
            if (B.type == LIN3) {
                sd.nl_slot = -1;
            } else {
                sd.nl_slot = nl;
                nl += (B.type == SOP && sd.role == PARENT) ? 6 : 3;  // d1 (and d2 for the two-term row kind)
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
            const double* W = M->seg[sgi].W;
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
        if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&M->d_markers, mk.size() * sizeof(DevMarker)) != cudaSuccess ||
            cudaMemcpy(M->d_markers, mk.data(), mk.size() * sizeof(DevMarker), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            delete M;
            return fail(BIONC_E_CUDA, "marker table upload failed");
        }
        M->nb_markers = d->nb_markers;
    }
    *out = M;
    return BIONC_OK;
}

int bionc_cuda_model_destroy(BioncModel* M) {
    if (M == nullptr) return BIONC_OK;
    cudaSetDevice(M->device);
    cudaFree(M->d_blob);
    cudaFree(M->d_markers);
    cudaFree(M->d_Q), cudaFree(M->d_Qd), cudaFree(M->d_Qdd), cudaFree(M->d_lam), cudaFree(M->d_inertia), cudaFree(M->d_h);
    cudaFree(M->d_status);
    if (M->stream != nullptr) cudaStreamDestroy(M->stream);
    for (auto& ps : M->pipe)
        if (ps != nullptr) cudaStreamDestroy(ps);
    delete M;
    return BIONC_OK;
}

int bionc_cuda_model_counts(const BioncModel* M, int32_t* nseg, int32_t* mj, int32_t* m) {
    if (M == nullptr) return fail(BIONC_E_INVALID, "null model");
    if (nseg) *nseg = M->hdr.N;
    if (mj) *mj = M->hdr.mj;
    if (m) *m = M->hdr.m;
    return BIONC_OK;
}

int bionc_cuda_model_layout(const BioncModel* M, int32_t* out) {
    if (M == nullptr || out == nullptr) return fail(BIONC_E_INVALID, "null argument");
    const DevHeader& h = M->hdr;
    const int32_t v[11] = {h.N, h.mj, h.nb, h.nslot, h.nslot2, h.nlevel, h.rnl, h.nzero, M->blob_valid ? h.spw : 0,
                           M->blob_valid ? (int32_t)M->smem_bytes : 0, h.serial_solve};
    std::memcpy(out, v, sizeof(v));
    return BIONC_OK;
}

int bionc_cuda_forward_dynamics(const BioncModel* Mc, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                double* Qdd, double* lam, int32_t* status, int64_t B, void* stream) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || B < 0) return fail(BIONC_E_INVALID, "null argument");
    if (B == 0) return BIONC_OK;
    std::lock_guard<std::mutex> lock(M->mu);
    DynOpts o;
    int rc = prepare(M, opt, o);
    if (rc != BIONC_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (status != nullptr) CUDA_TRY(cudaMemsetAsync(status, 0, B * sizeof(int32_t), st));
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    const unsigned grid = grid_for(M, B), block = 32 * M->hdr.N;
    const KernelSet ks = pick_kernels(needs_general(M, o), (int)block, M->hdr.spw);
    if ((rc = ensure_smem(ks.fd, M->smem_bytes)) != BIONC_OK) return rc;
    const unsigned char* blob = M->d_blob;
    int blob_bytes = M->blob_bytes;
    long long Bll = B;
    void* args[] = {&blob, &blob_bytes, &Q, &Qd, &inertia, &o, &Qdd, &lam, &status, &Bll};
    CUDA_TRY(cudaLaunchKernel(ks.fd, dim3(grid), dim3(block), args, M->smem_bytes, st));
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

int bionc_cuda_forward_dynamics_dense(const BioncModel* Mc, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                      const int64_t* select, int64_t count, double* Qdd, double* lam, int32_t* status,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || count < 0) return fail(BIONC_E_INVALID, "null argument");
    if (count == 0) return BIONC_OK;
    const size_t per = bionc_cuda_dense_workspace_bytes(M, 1);
    if (workspace == nullptr || workspace_bytes < per) return fail(BIONC_E_INVALID, "workspace smaller than bionc_cuda_dense_workspace_bytes(model, 1)");
    std::lock_guard<std::mutex> lock(M->mu);
    DynOpts o;
    int rc = prepare(M, opt, o);
    if (rc != BIONC_OK) return rc;
    long long slots = (long long)(workspace_bytes / per);
    if (slots > count) slots = count;
    if (slots > 148 * 8) slots = 148 * 8;
    EvalArgs ea{0, M->hdr.N, M->hdr.nblk, M->hdr.mj, M->hdr.m};
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    dense_kkt_kernel<<<(unsigned)slots, kDenseThreads, 0, (cudaStream_t)stream>>>(
        M->d_blob, ea, Q, Qd, inertia, o, reinterpret_cast<const long long*>(select), (long long)count, Qdd, lam, status,
        static_cast<double*>(workspace));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

int bionc_cuda_rk4(const BioncModel* Mc, double* Q, double* Qd, double h, const double* h_steps, int64_t n_steps,
                   int32_t normalize, const BioncDynOptions* opt, double* traj, int64_t traj_every, double* lam_last,
                   int32_t* status, int64_t B, void* stream) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    if (M == nullptr || Q == nullptr || Qd == nullptr || B < 0 || n_steps < 0) return fail(BIONC_E_INVALID, "null argument");
    if (traj != nullptr && traj_every < 1) return fail(BIONC_E_INVALID, "traj_every must be >= 1");
    if (B == 0 || n_steps == 0) return BIONC_OK;
    std::lock_guard<std::mutex> lock(M->mu);
    DynOpts o;
    int rc = prepare(M, opt, o);
    if (rc != BIONC_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (status != nullptr) CUDA_TRY(cudaMemsetAsync(status, 0, B * sizeof(int32_t), st));
    const double* inertia = opt != nullptr ? opt->inertia : nullptr;
    const unsigned grid = grid_for(M, B), block = 32 * M->hdr.N;
    if (traj == nullptr) traj_every = 1;
    const KernelSet ks = pick_kernels(needs_general(M, o), (int)block, M->hdr.spw);
    if ((rc = ensure_smem(ks.rk4, M->smem_bytes)) != BIONC_OK) return rc;
    const unsigned char* blob = M->d_blob;
    int blob_bytes = M->blob_bytes;
    long long Bll = B, nsteps = n_steps, tevery = traj_every;
    int norm = normalize;
    void* args[] = {&blob, &blob_bytes, &Q, &Qd, &inertia, &o, &h, &h_steps, &nsteps, &norm, &traj, &tevery, &lam_last, &status, &Bll};
    CUDA_TRY(cudaLaunchKernel(ks.rk4, dim3(grid), dim3(block), args, M->smem_bytes, st));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return BIONC_OK;
}

int bionc_cuda_eval(const BioncModel* Mc, int32_t which, const double* in, const BioncDynOptions* opt, double* out,
                    int64_t B, void* stream) {
    BioncModel* M = const_cast<BioncModel*>(Mc);
    if (M == nullptr || in == nullptr || out == nullptr || B < 0 || which < 0 || which > 12) return fail(BIONC_E_INVALID, "bad argument");
    if (B == 0) return BIONC_OK;
    std::lock_guard<std::mutex> lock(M->mu);
    DynOpts o;
    int rc = prepare(M, opt, o);
    if (rc != BIONC_OK) return rc;
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
    size_t out_elems = 0;
    switch (which) {
        case BIONC_EVAL_PHI_R: out_elems = (size_t)6 * N; break;
        case BIONC_EVAL_K_R: out_elems = (size_t)6 * N * n; break;
        case BIONC_EVAL_PHI_J: out_elems = mj; break;
        case BIONC_EVAL_K_J: out_elems = (size_t)mj * n; break;
        case BIONC_EVAL_PHI_H: case BIONC_EVAL_BIAS_H: out_elems = m; break;
        case BIONC_EVAL_K_H: out_elems = (size_t)m * n; break;
        case BIONC_EVAL_FEXT: out_elems = n; break;
    }
    if (out_elems == 0) return BIONC_OK;
    // vector outputs are written row by row in full; only the (sparse) Jacobians need the zero fill
    if (which == BIONC_EVAL_K_R || which == BIONC_EVAL_K_J || which == BIONC_EVAL_K_H)
        CUDA_TRY(cudaMemsetAsync(out, 0, out_elems * B * sizeof(double), st));
    EvalArgs ea{which, N, M->hdr.nblk, mj, m};
    const long long total = B * (long long)(N + M->hdr.nblk);
    const int block = 128;
    long long grid = (total + block - 1) / block;
    if (grid > 148LL * 64) grid = 148LL * 64;
    eval_kernel<<<(unsigned)grid, block, 0, st>>>(M->d_blob, ea, in, out, B);
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
static int ensure_host_scratch(BioncModel* M, long long B, long long nh, bool want_inertia) {
    CUDA_TRY(cudaSetDevice(M->device));
    if (M->stream == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking));
    for (auto& ps : M->pipe)
        if (ps == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    const size_t n = 12 * (size_t)M->hdr.N;
    if (B > M->cap_B) {
        cudaFree(M->d_Q), cudaFree(M->d_Qd), cudaFree(M->d_Qdd), cudaFree(M->d_lam), cudaFree(M->d_status), cudaFree(M->d_inertia);
        M->d_inertia = nullptr;
        CUDA_TRY(cudaMalloc(&M->d_Q, B * n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&M->d_Qd, B * n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&M->d_Qdd, B * n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&M->d_lam, B * (size_t)M->hdr.m * sizeof(double)));
        CUDA_TRY(cudaMalloc(&M->d_status, B * sizeof(int)));
        M->cap_B = B;
    }
    if (want_inertia && M->d_inertia == nullptr) CUDA_TRY(cudaMalloc(&M->d_inertia, M->cap_B * (size_t)M->hdr.N * 10 * sizeof(double)));
    if (nh > M->cap_h) {
        cudaFree(M->d_h);
        CUDA_TRY(cudaMalloc(&M->d_h, nh * sizeof(double)));
        M->cap_h = nh;
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

int bionc_cuda_forward_dynamics_host(BioncModel* M, const double* Q, const double* Qd, const BioncDynOptions* opt,
                                     double* Qdd, double* lam, int32_t* status, int64_t B) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || Qdd == nullptr || B < 0) return fail(BIONC_E_INVALID, "null argument");
    if (B == 0) return BIONC_OK;
    const bool per_sys = opt != nullptr && opt->inertia != nullptr;
    int rc = ensure_host_scratch(M, B, 0, per_sys);
    if (rc != BIONC_OK) return rc;
    const size_t n = 12 * (size_t)M->hdr.N, m = (size_t)M->hdr.m;
    const int64_t chunk = host_chunk(B);
    int k = 0;
    for (int64_t lo = 0; lo < B; lo += chunk, ++k) {
        const int64_t b = (B - lo < chunk) ? B - lo : chunk;
        cudaStream_t st = M->pipe[k % 3];
        const size_t nb = (size_t)b * n * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(M->d_Q + lo * n, Q + lo * n, nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(M->d_Qd + lo * n, Qd + lo * n, nb, cudaMemcpyHostToDevice, st));
        BioncDynOptions o2 = opt != nullptr ? *opt : BioncDynOptions{};
        if (per_sys) {
            const size_t ni = (size_t)M->hdr.N * 10;
            CUDA_TRY(cudaMemcpyAsync(M->d_inertia + lo * ni, opt->inertia + lo * ni, (size_t)b * ni * sizeof(double), cudaMemcpyHostToDevice, st));
            o2.inertia = M->d_inertia + lo * ni;
        }
        rc = bionc_cuda_forward_dynamics(M, M->d_Q + lo * n, M->d_Qd + lo * n, &o2, M->d_Qdd + lo * n,
                                         lam != nullptr ? M->d_lam + lo * m : nullptr, status != nullptr ? M->d_status + lo : nullptr, b, st);
        if (rc != BIONC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(Qdd + lo * n, M->d_Qdd + lo * n, nb, cudaMemcpyDeviceToHost, st));
        if (lam != nullptr) CUDA_TRY(cudaMemcpyAsync(lam + lo * m, M->d_lam + lo * m, (size_t)b * m * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (status != nullptr) CUDA_TRY(cudaMemcpyAsync(status + lo, M->d_status + lo, (size_t)b * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    for (auto& ps : M->pipe) CUDA_TRY(cudaStreamSynchronize(ps));
    return BIONC_OK;
}

int bionc_cuda_rk4_host(BioncModel* M, double* Q, double* Qd, double h, const double* h_steps, int64_t n_steps,
                        int32_t normalize, const BioncDynOptions* opt, int32_t* status, int64_t B) {
    if (M == nullptr || Q == nullptr || Qd == nullptr || B < 0 || n_steps < 0) return fail(BIONC_E_INVALID, "null argument");
    if (B == 0 || n_steps == 0) return BIONC_OK;
    const bool per_sys = opt != nullptr && opt->inertia != nullptr;
    int rc = ensure_host_scratch(M, B, h_steps != nullptr ? n_steps : 0, per_sys);
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
        CUDA_TRY(cudaMemcpyAsync(M->d_Q + lo * n, Q + lo * n, nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(M->d_Qd + lo * n, Qd + lo * n, nb, cudaMemcpyHostToDevice, st));
        BioncDynOptions o2 = opt != nullptr ? *opt : BioncDynOptions{};
        if (per_sys) {
            const size_t ni = (size_t)M->hdr.N * 10;
            CUDA_TRY(cudaMemcpyAsync(M->d_inertia + lo * ni, opt->inertia + lo * ni, (size_t)b * ni * sizeof(double), cudaMemcpyHostToDevice, st));
            o2.inertia = M->d_inertia + lo * ni;
        }
        rc = bionc_cuda_rk4(M, M->d_Q + lo * n, M->d_Qd + lo * n, h, h_steps != nullptr ? M->d_h : nullptr, n_steps, normalize,
                            &o2, nullptr, 1, nullptr, status != nullptr ? M->d_status + lo : nullptr, b, st);
        if (rc != BIONC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(Q + lo * n, M->d_Q + lo * n, nb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(Qd + lo * n, M->d_Qd + lo * n, nb, cudaMemcpyDeviceToHost, st));
        if (status != nullptr) CUDA_TRY(cudaMemcpyAsync(status + lo, M->d_status + lo, (size_t)b * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    for (auto& ps : M->pipe) CUDA_TRY(cudaStreamSynchronize(ps));
    return BIONC_OK;
}

// FP64 roofline denominator: a DFMA-only kernel (8 independent accumulator chains per thread,
// 2048 threads per SM resident) timed with CUDA events; MEASURED_PEAKS.json has no FP64 figure.
int bionc_cuda_measure_fp64_peak(int device, int repeats, double* tflops_best, double* tflops_sustained) {
    if (tflops_best == nullptr) return fail(BIONC_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
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

int64_t bionc_cuda_launch_count(void) { return g_launches.load(); }
const char* bionc_cuda_last_error(void) { return g_last_error.c_str(); }
const char* bionc_cuda_version(void) { return "bionc_cuda 0.1.0 (sm_100a, fp64)"; }

}  // extern "C"
