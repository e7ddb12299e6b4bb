// Device-side model table ("blob") shared by every kernel.
//
// The host lowers a BioncModelDesc (include/bionc_cuda.h) to one contiguous, immutable blob:
//   DevHeader | DevSeg[N] | DevBlk[nblk] | DevSide[nside] | symbolic tables
// which each CTA copies into shared memory once.  Everything a lane needs about "its" segment and
// the constraint blocks touching it is found through DevSeg::side_begin/side_end.  External forces are
// not part of the model: a call that has them passes its own table (DevFextTab | DevFext[nf]), copied
// to shared memory behind the blob.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

namespace bionc {

constexpr int kMaxSegments = 32;  // one warp per segment, at most 1024 threads per CTA
constexpr int kBlkNParam = 16;
constexpr int kFextNParam = 24;

enum BlkType : int { LIN3 = 0, DOT = 1, CLEN = 2, SOP = 3 };
enum Role : int { CHILD = 0, PARENT = 1 };

struct DevHeader {
    int N;        // segments
    int nblk;     // constraint blocks
    int nside;    // (block, segment) incidences
    int mj;       // joint-constraint rows
    int nb;       // 3x3 block rows ("nodes") of the joint-level Schur system (mj padded up to a multiple of 3)
    int m;        // holonomic rows = mj + 6 N
    int spw;      // systems per CTA without external forces: 32 (one lane per system), or 16 when the working set of 32 does not fit
    int rnl;      // doubles of shared memory per lane for the 6-vectors of the non-LIN3 rows touching a segment
    int pad0;
    int blob_bytes;
    int off_seg, off_blk, off_side, pad1;  // byte offsets inside the blob
    // symbolic factorisation of the joint-level Schur complement (computed on the host, see capi.cu)
    int nslot;    // stored 3x3 blocks of the lower triangle: structural non-zeros + fill
    int nslot2;   // blocks that can receive contributions from two segments (second, exclusive-writer copy)
    int nzero;    // entries of the zero list (slots that are not fully overwritten by the assembly)
    int ncomb;    // entries of the combine list (slot, slot2) pairs
    int nunit;    // entries of the unit list: (diagonal element, row) of every padding row
    int ncol;     // total length of the column lists
    int off_sym;  // byte offset of the symbolic tables (ints): slot_tab[nb*nb] slot2_tab[nb*nb] zero[nzero]
                  // comb[2*ncomb] unit[2*nunit] colptr[nb+1] colrow[ncol] lvlptr[nlevel+1] lvlpiv[nb]
                  // owner[nb] (warp of block row I) pivmask[nb] (warps taking part in pivot K)
                  // pivwriter[nb] (warp that stores the inverse pivot and y-hat of pivot K)
                  // nscale[nb] (float bits: nominal diagonal scale of each node)
    int nlevel;   // levels of the elimination tree: pivots of one level are mutually independent
    int tmem_slot;     // scratch word of the shared-memory copy: TMEM base address of the CTA
    int serial_solve;  // joint-level schedule: 0 level-parallel, 1 serial on the first warp, 2 two leaves + root on two warps
};

struct DevSeg {
    double kc[12];   // what the dynamics need of (m, c, J): N3 = J - m c c^T (xx xy xz yy yz zz), c (3), m, 1/m, pad
    double fg[4];    // gravity: natural force on slot s is fg[s] * e_z   (natural_segment.py:562-572)
    double phic[4];  // L cos(gamma), cos(beta), L^2, L cos(alpha)         (natural_segment.py:441-446)
    double mcj[10];  // mass, c(3), J00 J01 J02 J11 J12 J22 (for bionc_cuda_mass_matrix / documentation)
    int side_begin, side_end;
    int rigid_row;  // first holonomic row of the 6 rigid-body rows
    int pad0, pad1, pad2, pad3, pad4;
};

struct DevBlk {
    double p[kBlkNParam];
    int type, parent, child;
    int jrow;   // first row inside the joint-level Schur system (holonomic order without rigid rows)
    int row_h;  // first row, holonomic order (lambda / K / Phi / bias outputs)
    int row_j;  // first row, joint-index order (joint_constraints*)
    int pad0, pad1;
};

// One (block, segment) incidence with everything the per-segment phases need resolved on the host.
struct DevSide {
    int blk;
    int role;      // CHILD / PARENT
    int nl_slot;   // offset (in doubles) of the row's 6-vector in the lane's storage, -1 for LIN3
    int node;      // 3x3 node of the joint-level system holding the block's first row (jrow / 3)
    int rhs_off;   // row of this side's copy of the right-hand side inside [r0 | r1]: jrow (+ 3 nb for the parent copy)
    int lam_off;   // jrow: first row of the block in lambda_j
    int row_h;     // first row, holonomic order
    int ground;    // bit 0: the block has no parent segment (the child also settles the second copy of the right-hand side);
                   // bit 1: non-LIN3 row whose force direction f_dir is not identically zero (6 stored doubles instead of 3)
    // point-to-point (LIN3) sides: the signed coefficient vector as seen from the centre of mass of the MODEL's inertial
    // parameters, lev = (sigma, r[3]), sigma = a_rp + a_rd, r = (a_u, -a_rd, a_w) - sigma c (see Lever, dynamics.cuh)
    double lev[4];
};

struct DevFext {
    double p[kFextNParam];  // torque(3) force(3) point(3) Tinv(9, row-major)
    int segment, kind;
    int slot;   // position in the caller's force list: per-system values are (B, F, 6) in that order
    int other;  // kind 4 (reaction of a joint force): the segment the force itself acts on (the joint's child)
};

// header of a call's force table; the forces follow it, sorted by segment
struct DevFextTab {
    int nf, pad0, pad1, pad2;
    int begin[kMaxSegments + 4];  // forces of segment s: begin[s] .. begin[s + 1]
};

// external forces of one launch (kernel parameter)
struct FextArgs {
    const unsigned char* tab;  // device: DevFextTab | DevFext[nf], or nullptr
    int tab_bytes;             // multiple of 16
    int nf;
    const double* vals;        // device, optional: per-system (torque, force), (steps, B, nf, 6)
    long long step_stride;     // doubles between the slices of two RK4 steps (0: one slice for every step)
    long long steps;           // slices
};

// Shared-memory working set of one CTA (spc systems, one warp per segment), in doubles (layout: dynamics.cuh).
constexpr int kSegConsts = 11;  // N3 (6), natural centre of mass (3), m, 1 / m
__host__ __device__ inline int cta_smem_doubles(int spc, int N, int nb, int nslot, int nslot2, int rnl, bool persys = false) {
    return spc * (2 * 12 * N              // published Q, Qdot of the current stage
                 + (persys ? kSegConsts * N : 0)  // per-system segment constants (per-system inertial parameters only)
                 + 9 * (nslot + nslot2)  // stored 3x3 blocks of the joint-level Schur matrix (+ second copies)
                 + 3 * 3 * nb            // rhs copy 0 / copy 1 (-> y-hat), lambda_j
                 + 6 * nb                // inverse pivot blocks
                 + rnl * N);             // d-vectors of the non-LIN3 rows (rnl doubles per segment)
}

}  // namespace bionc
