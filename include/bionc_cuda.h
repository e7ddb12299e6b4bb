/*
 * bionc_cuda.h -- C ABI of the B200 (sm_100a, FP64) constrained forward-dynamics backend.
 *
 * The reference (Ipuch/bioNC) is pure Python and has no FFI of its own; the boundary this library
 * sits behind is the backend-subpackage convention (bionc/bionc_numpy/__init__.py:1-27).  Each
 * entry point below names the reference method it replaces.  All pointers in the "device" entry
 * points are CUDA device pointers owned by the caller (PyTorch tensors on the Python side); the
 * library allocates only the immutable model table.  The "_host" entry points take host pointers
 * and do the host<->device copies themselves (the end-to-end path a non-PyTorch caller binds).
 *
 * Every function returns 0 on success or a negative BIONC_E_* code; no exception crosses the ABI.
 * Arrays are FP64, C-contiguous, batch-major: Q, Qdot, Qddot (B, 12 N); lambda (B, m) in the
 * holonomic row order of bionc/bionc_numpy/biomechanical_model.py:154-258.
 */
#ifndef BIONC_CUDA_H
#define BIONC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIONC_OK 0
#define BIONC_E_INVALID -1   /* bad argument / malformed model description */
#define BIONC_E_CUDA -2      /* a CUDA runtime call failed (see bionc_cuda_last_error) */
#define BIONC_E_TOO_LARGE -3 /* model does not fit the per-SM shared memory budget of the kernels */

/* primitive constraint-block kinds (bionc_b200/bionc_cuda/model_table.py) */
#define BIONC_BLK_LIN3 0
#define BIONC_BLK_DOT 1
#define BIONC_BLK_CLEN 2
#define BIONC_BLK_SOP 3
#define BIONC_BLK_NPARAM 16

/* external-force kinds (bionc/bionc_numpy/external_force*.py) */
#define BIONC_FEXT_GLOBAL_ON_PROXIMAL 0
#define BIONC_FEXT_GLOBAL_LOCAL_POINT 1
#define BIONC_FEXT_GLOBAL 2
#define BIONC_FEXT_LOCAL 3
#define BIONC_FEXT_JOINT_REACTION 4 /* reaction of a joint generalized force on the joint's parent (generalized_force.py:98-129):
                                       `segment` is the parent, param[18] the index of the child the force (torque, force) itself
                                       acts on; the parent receives -(torque + (rp_parent - rp_child) x force, force) */
#define BIONC_FEXT_NPARAM 24 /* torque(3) force(3) point(3) Tinv(9) other segment(1, kind 4) pad(5) */

/* per-system status bits written by forward_dynamics / rk4 */
#define BIONC_STATUS_OK 0
#define BIONC_STATUS_SINGULAR 1  /* zero or non-finite pivot (reference: numpy.linalg.LinAlgError) */
#define BIONC_STATUS_NONFINITE 2 /* non-finite state or acceleration */
#define BIONC_STATUS_SMALL_PIVOT 4 /* a pivot below 1e-11 of the diagonal scale: (near-)singular constraints, the
                                      unpivoted LDL^T result may be inaccurate (e.g. redundant constraints) */
#define BIONC_STATUS_PIVOTED 8 /* the values of this system come from bionc_cuda_forward_dynamics_dense */

/* which quantity bionc_cuda_eval computes */
#define BIONC_EVAL_PHI_R 0  /* rigid_body_constraints(Q)                 -> (B, 6N)      biomechanical_model_segments.py:26-42 */
#define BIONC_EVAL_K_R 1    /* rigid_body_constraints_jacobian(Q)        -> (B, 6N, 12N) biomechanical_model_segments.py:62-79 */
#define BIONC_EVAL_PHI_J 2  /* joint_constraints(Q)                      -> (B, mj)      biomechanical_model_joints.py:13-41 */
#define BIONC_EVAL_K_J 3    /* joint_constraints_jacobian(Q)             -> (B, mj, 12N) biomechanical_model_joints.py:43-77 */
#define BIONC_EVAL_PHI_H 4  /* holonomic_constraints(Q)                  -> (B, m)       biomechanical_model.py:154-197 */
#define BIONC_EVAL_K_H 5    /* holonomic_constraints_jacobian(Q)         -> (B, m, 12N)  biomechanical_model.py:199-258 */
#define BIONC_EVAL_BIAS_H 6 /* holonomic_constraints_acceleration_bias(Qdot) -> (B, m)   biomechanical_model.py:260-317 */
#define BIONC_EVAL_FEXT 7   /* ExternalForceSet.to_natural_external_forces(Q) -> (B, 12N) external_force.py:143-166 */
/* post-processing: what users evaluate frame by frame after an integration */
#define BIONC_EVAL_COM 8        /* center_of_mass_position(Q)  -> (B, 3, N)          biomechanical_model_markers.py:59-79 */
#define BIONC_EVAL_MARKERS 9    /* markers(Q), forward kinematics -> (B, 3, nb_markers) biomechanical_model_markers.py:32-57 */
#define BIONC_EVAL_KINETIC 10   /* kinetic_energy(Qdot)        -> (B)                biomechanical_model.py:98-113 */
#define BIONC_EVAL_POTENTIAL 11 /* potential_energy(Q)         -> (B)                biomechanical_model.py:115-133 */
#define BIONC_EVAL_RESIDUALS 12 /* [|Phi_r(Q)|_2, |Phi_j(Q)|_2] -> (B, 2)            bionc_numpy/time_series_utils.py:6-41 */

/* Flat model description: the arrays of a compiled ModelTable (host pointers, copied at create). */
typedef struct BioncModelDesc {
    int32_t nb_segments;         /* N */
    int32_t nb_blocks;           /* constraint blocks */
    const double* seg_phi_const; /* (N, 4)  [L cos(gamma), cos(beta), L^2, L cos(alpha)]  natural_segment.py:441-446 */
    const double* seg_mass;      /* (N)                                                                             */
    const double* seg_com;       /* (N, 3)  natural centre of mass                                                  */
    const double* seg_J;         /* (N, 9)  natural pseudo-inertia   natural_inertial_parameters.py:153-177         */
    const int32_t* seg_rigid_row;/* (N)     first holonomic row of the segment's 6 rigid-body rows                  */
    const int32_t* blk_type;     /* (nblk)  BIONC_BLK_*                                                              */
    const int32_t* blk_parent;   /* (nblk)  -1 = ground                                                              */
    const int32_t* blk_child;    /* (nblk)                                                                           */
    const int32_t* blk_row_h;    /* (nblk)  first row, holonomic order                                               */
    const int32_t* blk_row_j;    /* (nblk)  first row, joint-index order                                             */
    const double* blk_param;     /* (nblk, BIONC_BLK_NPARAM)                                                         */
    int32_t nb_markers;          /* markers in model order (0 = none), natural_marker.py:39-71                       */
    const int32_t* marker_segment; /* (nb_markers) segment index                                                     */
    const double* marker_position; /* (nb_markers, 3) natural coordinates (n_u, n_v, n_w)                            */
} BioncModelDesc;

/* External forces of one call (ExternalForceSet, external_force.py:13-180).  The list (segment, kind, application
 * point, local frame) is shared by the batch; the (torque, force) values are either shared too (those in `param`) or
 * given per system -- the reference evaluates ExternalForceSet.to_natural_external_forces(Q) for every call of its
 * dynamics closure (external_force.py:143-180, utils/ode_solver.py:107-116), so every trajectory may carry its own
 * loads and change them from step to step.  Nothing of this is kept in the model: the description is read during the
 * call and may be freed or changed as soon as the call returns. */
typedef struct BioncFextDesc {
    int32_t nb_forces;
    const int32_t* segment; /* (F) host */
    const int32_t* kind;    /* (F) host, BIONC_FEXT_* */
    const double* param;    /* (F, BIONC_FEXT_NPARAM) host, torque first (external_force.py:84-85) */
    /* optional per-system values replacing torque(3), force(3) of `param`: (values_steps, B, F, 6), forces in the order
     * of this list.  DEVICE pointer for the device entry points, HOST pointer for the _host ones; NULL = shared values.
     * values_steps <= 1: one slice for the whole call.  values_steps > 1 (rk4 only): step k of the launch uses slice
     * min(k, values_steps - 1), held over the four stages of the step. */
    const double* values;
    int64_t values_steps;
} BioncFextDesc;

/* Options common to forward_dynamics and rk4. */
typedef struct BioncDynOptions {
    int32_t use_stabilization; /* Baumgarte, biomechanical_model.py:416-421 */
    double alpha, beta;
    const BioncFextDesc* fext; /* NULL = none (uploaded per call on the caller's stream, a few hundred bytes) */
    /* optional per-system natural inertial parameters (device for the device entry points, host
     * for the _host ones): (B, N, 10) = [mass, c0, c1, c2, J00, J01, J02, J11, J12, J22]; NULL = use
     * the model's own (config 5: "perturbed inertial parameters") */
    const double* inertia;
} BioncDynOptions;

typedef struct BioncModel BioncModel;

/* replaces BiomechanicalModel construction / to_mx() (bionc_numpy/biomechanical_model.py:43-75) */
int bionc_cuda_model_create(const BioncModelDesc* desc, int device, BioncModel** out);
int bionc_cuda_model_destroy(BioncModel* model);
/* counts of protocols/biomechanical_model.py:347-453 */
int bionc_cuda_model_counts(const BioncModel* model, int32_t* nb_segments, int32_t* nb_joint_constraints,
                            int32_t* nb_holonomic_constraints);

/* Model-specialised integrator (optional; the counterpart of the reference's model.to_mx() code generation,
 * bionc_casadi/biomechanical_model.py, for this path).  Compiles, with NVRTC for sm_100a, a build of the fused RK4 kernel
 * in which this model's tables are compile-time constants, for calls shaped like `options` (stabilisation on / off,
 * external forces present, per-system inertial parameters present: only the presence of fext / inertia is read), and --
 * load != 0 -- loads it on the model's device.  bionc_cuda_rk4 / _rk4_host calls of that shape then launch it instead
 * of the generic kernel; every other call is unaffected.  Seconds of host time per shape, once; results agree with the
 * generic kernel to rounding.  systems_per_cta: 0 = the device's launch plan (needs the device), or 16 / 32 (load == 0:
 * compile only, e.g. on a machine without a GPU).  Without libnvrtc / libcuda at run time: BIONC_E_CUDA, the generic
 * kernels remain. */
int bionc_cuda_model_specialize(BioncModel* model, const BioncDynOptions* options, int32_t systems_per_cta, int32_t load);
/* the generated CUDA source of that kernel (diagnostics); needed: bytes including the terminator */
int bionc_cuda_model_spec_source(const BioncModel* model, const BioncDynOptions* options, int32_t systems_per_cta, char* buf,
                                 size_t cap, size_t* needed);
/* number of loaded model-specialised kernels; cubin_bytes (optional): total size of the compiled images */
int bionc_cuda_model_specialized(const BioncModel* model, int64_t* cubin_bytes);

/* diagnostics: how the model was laid out for the kernels.  out[11] = {segments, joint rows, 3x3 nodes of the
 * joint-level system, stored blocks, second copies, elimination-tree levels, doubles of per-segment row
 * storage, zero-list entries, systems per CTA, shared-memory bytes per CTA, serial joint-level schedule (0/1)};
 * systems and bytes per CTA are 0 until the first launch has sized the CTA for the device */
int bionc_cuda_model_layout(const BioncModel* model, int32_t* out);
/* diagnostics: first row of every constraint block (in the order of BioncModelDesc::blk_*) inside the joint-level
 * system as the kernels order it: node = row / 3 is eliminated in ascending order (fill-reducing ordering of the host
 * symbolic analysis); lets a host program replay the block elimination.  out: nb_blocks values */
int bionc_cuda_model_joint_rows(const BioncModel* model, int32_t* first_row_of_block);

/* Device entry points: the per-system arrays (Q, Qdot, Qddot, traj, the in / out of bionc_cuda_eval) are read and
 * written with 16-byte vector accesses and must be 16-byte aligned (cudaMalloc and every row of a (B, 12 N) array are;
 * a misaligned pointer returns BIONC_E_INVALID, no launch).  The _host entry points take any alignment. */

/* replaces BiomechanicalModel.forward_dynamics (bionc_numpy/biomechanical_model.py:351-429), batched.
 * lambda and status may be NULL. */
int bionc_cuda_forward_dynamics(const BioncModel* model, const double* Q, const double* Qdot,
                                const BioncDynOptions* opt, double* Qddot, double* lambda, int32_t* status,
                                int64_t B, void* stream);

/* Pivoted variant = the reference's own algorithm on the device: the dense augmented matrix [G K^T; K 0] of
 * BiomechanicalModel.augmented_mass_matrix (bionc_numpy/biomechanical_model.py:335-349) solved by LU with
 * partial pivoting like numpy.linalg.solve -> LAPACK dgesv (biomechanical_model.py:423-426), one CTA per
 * system.  Meant for the systems the fast path flags (indefinite mass matrices whose unpivoted LDL^T meets a
 * small pivot); `select` lists `count` system indices (NULL -> systems 0..count-1).  Rows of Qddot / lambda /
 * status that belong to the selected systems are overwritten; status becomes BIONC_STATUS_PIVOTED, plus
 * SINGULAR for an exactly singular matrix (numpy: LinAlgError) or NONFINITE.  `workspace` is caller-owned
 * device memory of at least bionc_cuda_dense_workspace_bytes(model, 1) bytes; more lets more systems run
 * concurrently (bionc_cuda_dense_workspace_bytes(model, count) = enough for full concurrency). */
size_t bionc_cuda_dense_workspace_bytes(const BioncModel* model, int64_t count);
int bionc_cuda_forward_dynamics_dense(const BioncModel* model, const double* Q, const double* Qdot,
                                      const BioncDynOptions* opt, const int64_t* select, int64_t count,
                                      double* Qddot, double* lambda, int32_t* status, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* replaces RK4 + the dynamics closure of forward_integration (utils/ode_solver.py:8-53, :107-120),
 * all four stages fused, n_steps steps per launch, in place on Q / Qdot.
 *   h_steps : device array of n_steps step sizes (h = t[i+1]-t[i], ode_solver.py:41) or NULL -> constant h
 *   normalize: post-step renormalisation of u, w (model.normalized_coordinates, ode_solver.py:50-52)
 *   traj / traj_every: if traj != NULL the state [Q; Qdot] is also written after every
 *       traj_every-th step to traj[(step/traj_every - 1), b, 0:24N]  (n_steps / traj_every snapshots)
 *   lambda_last: (B, m) multipliers of the first stage of the last step, or NULL */
int bionc_cuda_rk4(const BioncModel* model, double* Q, double* Qdot, double h, const double* h_steps,
                   int64_t n_steps, int32_t normalize, const BioncDynOptions* opt, double* traj,
                   int64_t traj_every, double* lambda_last, int32_t* status, int64_t B, void* stream);

/* component evaluators (the mass_matrix / *_constraints* / *_jacobian API); `in` is Q or Qdot */
int bionc_cuda_eval(const BioncModel* model, int32_t which, const double* in, const BioncDynOptions* opt,
                    double* out, int64_t B, void* stream);
/* BiomechanicalModel.mass_matrix (protocols/biomechanical_model.py:455-466): host output (12N, 12N) */
int bionc_cuda_mass_matrix(const BioncModel* model, double* G_host);

/* host-buffer variants: H2D, kernel, D2H on an internal stream, synchronous on return */
int bionc_cuda_forward_dynamics_host(BioncModel* model, const double* Q, const double* Qdot,
                                     const BioncDynOptions* opt, double* Qddot, double* lambda, int32_t* status,
                                     int64_t B);
int bionc_cuda_rk4_host(BioncModel* model, double* Q, double* Qdot, double h, const double* h_steps,
                        int64_t n_steps, int32_t normalize, const BioncDynOptions* opt, int32_t* status, int64_t B);

/* FP64 roofline denominator: DFMA-only micro-benchmark timed with CUDA events (best single launch and
 * the average over `repeats` back-to-back launches), in TFLOP/s */
int bionc_cuda_measure_fp64_peak(int device, int repeats, double* tflops_best, double* tflops_sustained);

/* debug aid: the dynamics kernels fill the selected regions of their shared-memory working set with NaN before the first
 * evaluation (bit 0 S', 1 second copies, 2 / 3 right-hand-side copies, 4 lambda_j, 5 inverse pivots, 6 row vectors of the
 * non-LIN3 blocks).  A correct kernel writes every word before reading it, so no result may change; 0 switches it off.
 * Bit 8 makes bionc_cuda_forward_dynamics use its per-lane load/store kernel where it would pick the bulk-copy one
 * (forward_dynamics_rows_kernel), so that tests and measurements can compare the two. */
int bionc_cuda_debug_poison(int32_t mask);

/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t bionc_cuda_launch_count(void);
const char* bionc_cuda_last_error(void);
const char* bionc_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BIONC_CUDA_H */
