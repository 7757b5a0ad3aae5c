// kparams.h -- plain structs passed BY VALUE to the kernels (they live in the constant bank of the
// kernel-parameter space, so warp-uniform reads are broadcast LDCs).  Filled by the host side in
// c_abi.cu from the gpmp2b_* descriptors.
#pragma once
#include <stdint.h>

// -DGPMP2B_DEBUG_BOUNDS: in-tree memory-safety net (compute-sanitizer is not available on the GPU pool).  Every checked
// index that leaves its array prints the site and traps, which surfaces as a CUDA error of the call (libgpmp2b_dbg.so,
// `make debug`; tests/test_debug_bounds.py).  Checked: every SDF cell index against the field size (all kernels), every
// shared-memory store of the tensor-core solve kernel and of its assembly, and the per-trajectory scratch indices of the
// phase kernels.  Compiles to nothing in the release library.
#ifdef GPMP2B_DEBUG_BOUNDS
#include <cstdio>
#define DBG_IDX(idx, n, what)                                                                                     \
  do {                                                                                                            \
    if ((long long)(idx) < 0 || (long long)(idx) >= (long long)(n)) {                                             \
      printf("GPMP2B_DEBUG_BOUNDS: %s index %lld outside [0, %lld) at %s:%d (block %d thread %d)\n", what,         \
             (long long)(idx), (long long)(n), __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x);             \
      __trap();                                                                                                   \
    }                                                                                                             \
  } while (0)
#else
#define DBG_IDX(idx, n, what) ((void)0)
#endif

#define KP_MAX_DOF 7        // system dof the kernels are instantiated for (T = D(D+1)/2 <= 32 lanes)
#define KP_MAX_JOINTS 7
#define KP_MAX_SPHERES 64
#define KP_MAX_INTER 24     // obs_check_inter
#define KP_MAX_SELF_PAIRS 32

struct KRobot {
  int32_t kind, arm_dof, dof, n_spheres;
  double base[12];                 // 3x4 row-major [R|t]: ARM base pose / MOBILE base_T_arm | base_T_arm1 | base_T_torso
  // the other Pose2Vector robots (gpmp2b.h: GPMP2B_ROBOT_POSE2_MOBILE_2ARMS / _VETLIN_ARM / _VETLIN_2ARMS).  arm_dof counts
  // the DH joints of both arms; n1 = joints of arm 1; nb = base coordinates in front of them (0 arm, 3 Pose2, 4 with the
  // linear actuator); lift = 0 / +1 / -1 (reverse_linact); base2 = base_T_arm2 | torso_T_arm | torso_T_arm1; base3 = torso_T_arm2
  int32_t n1, nb, lift, pad_;
  double base2[12], base3[12];
  double ca[KP_MAX_JOINTS], sa[KP_MAX_JOINTS];   // cos/sin(alpha_j)
  double a[KP_MAX_JOINTS], d[KP_MAX_JOINTS], bias[KP_MAX_JOINTS];
  // spheres sorted by link id (host keeps the permutation for debug outputs)
  int32_t sph_link[KP_MAX_SPHERES];
  int32_t sph_begin[KP_MAX_JOINTS + 3];   // spheres of link l are [sph_begin[l], sph_begin[l+1])
  int32_t sph_orig[KP_MAX_SPHERES];   // original index of sorted sphere s
  double sph_r[KP_MAX_SPHERES];
  double sph_c[KP_MAX_SPHERES][3];
};

struct KSdf {
  int32_t ndim, rows, cols, nz;
  double ox, oy, oz;         // origin
  double hx, hy, hz;         // inclusive upper bound: origin + (n-1)*cell  (SignedDistanceField.h:105-107)
  double cell, inv_cell;
  // "quad" layout: cell (z, col, row) holds {v[r][c], v[r+1][c], v[r][c+1], v[r+1][c+1]} of slice z, 32 bytes =
  // one L2 sector, fetched with ONE 256-bit load (indices clamped at the upper edges).  A trilinear lookup is two
  // sector requests instead of eight 8-byte ones: the kernel was bound by the SM's ~1 divergent request per clock.
  const double* quad;        // [z][col][row][4]
};

struct KSetting {
  int32_t D, N, K, opt_type, max_iter, flag_pos_limit, flag_vel_limit, qc_identity;
  double rel_thresh;
  double epsilon;
  double inv_cost_sigma;              // 1 / cost_sigma
  double conf_prior_w, vel_prior_w;   // 1 / sigma^2
  // optional workspace goal on x_T (gpmp2b_setting.goal_*): GoalFactorArm / GaussianPriorWorkspacePositionArm
  double end_conf_prior_w;            // conf_prior_w, or 0 when the goal factor replaces PriorFactor(x_T, end_conf)
  double goal_w;                      // 1 / goal_sigma^2
  double goal_pos[3];
  int32_t goal_enabled, goal_link;    // goal_enabled: 1 position goal, 2 pose goal; goal_link: 0-based link frame (resolved: -1 -> last)
  // optional self-collision pairs checked at every support state (gpmp2b_setting.self_collision_data): SelfCollisionArm
  int32_t n_self, pad_;
  int32_t self_a[KP_MAX_SELF_PAIRS], self_b[KP_MAX_SELF_PAIRS];   // sphere indices in the kernel's (link-sorted) order
  double self_eps[KP_MAX_SELF_PAIRS];    // r_A + r_B + epsilon
  double self_isig[KP_MAX_SELF_PAIRS];   // 1 / sigma
  // optional VehicleDynamicsFactorPose2Vector on every support state (Pose2MobileArm): e = v_i(1); 1 / sigma^2 or 0
  double veh_w;
  // optional GaussianPriorWorkspaceOrientation on support states orient_first..orient_last
  int32_t orient_enabled, orient_link, orient_first, orient_last;
  double orient_w;        // 1 / sigma^2
  double orient_R[9];     // desired rotation, row-major
  double goal_R[9];       // goal_enabled == 2 (GaussianPriorWorkspacePose): desired rotation, row-major; goal_pos = translation
  // optional PriorFactor pair on support state fix_index (gpmp2b_setting.fix_*: ISAM2TrajOptimizer::fixConfigAndVel), weights
  // conf_prior_w / vel_prior_w; the targets are per problem (KProblem.fix_conf_pp / fix_vel_pp)
  int32_t fix_enabled, fix_index;
  double delta_t;
  // GP prior (GaussianProcessPriorLinear): Q^-1 = qi (x) Qc^-1, Hessian blocks s11 = Phi^T qi Phi,
  // s12 = -Phi^T qi, s22 = qi, all 2x2 scalar matrices to be Kronecker-multiplied by Qc^-1
  double qi[2][2], s11[2][2], s12[2][2], s22[2][2];
  double Qc_inv[KP_MAX_DOF * KP_MAX_DOF];     // row-major D x D
  // GP interpolation weights per interior point j=1..K: x(tau) = w0 x1 + w1 v1 + w2 x2 + w3 v2
  double gpw[KP_MAX_INTER][4];
  // products of the above, order: w0w0 w0w1 w1w1 | w0w2 w0w3 w1w2 w1w3 | w2w2 w2w3 w3w3
  double gpww[KP_MAX_INTER][10];
  // limit factors (JointLimitFactorVector / VelocityLimitFactorVector)
  double pos_lo[KP_MAX_DOF], pos_hi[KP_MAX_DOF], pos_th[KP_MAX_DOF], pos_w[KP_MAX_DOF];   // w = 1/sigma^2
  double vel_lim[KP_MAX_DOF], vel_th[KP_MAX_DOF], vel_w[KP_MAX_DOF];
};

// per-call problem pointers (device memory)
struct KProblem {
  int64_t B;
  const double* start_conf;   // [B][D]
  const double* start_vel;
  const double* end_conf;
  const double* end_vel;
  const double* init_traj;    // [B][2*N*D]
  double* out_traj;           // [B][2*N*D]
  double* out_error;          // [B] or null
  double* out_coll_cost;      // [B] or null
  int32_t* out_iters;         // [B] or null
  int32_t* out_status;        // [B] or null
  // debug outputs (linearize mode)
  double* out_Hdiag;          // [B][N][b][b]
  double* out_Hoff;           // [B][N-1][b][b]
  double* out_g;              // [B][N][b]
  // debug outputs (obstacle-errors mode)
  double* out_obs_err;        // [B][C][S]
  double* out_centers;        // [B][C][S][3] or null
  // optional per-problem workspace targets (gpmp2b_setting.*_batch): null = the setting's shared value
  const double* goal_pos_pp;  // [B][3]
  const double* goal_R_pp;    // [B][9]
  const double* orient_R_pp;  // [B][9]
  const double* fix_conf_pp;  // [B][D]  (st.fix_enabled)
  const double* fix_vel_pp;   // [B][D]
  // scratch
  double* h_backup;           // [grid][hsize] backup of H for lambda retries
  unsigned long long* counters;  // [3] linearizations, solves, error evals (summed over batch)
  unsigned long long* queue;     // dynamic work queue of this launch: next problem index minus gridDim.x (zeroed by the host)
  // ---- phase-kernel pipeline (pk_kernels.cuh): per-trajectory state between the kernels of a round ----
  double* pk_state;              // [B][pk_state_size]: xs[N*b] (state-major) | dl[N*b] | 16 scalars (enum PkScalar)
  double* pk_mlist;              // [B][C][RS]: per-configuration (M, cv) of the last linearization, configuration-major
  int32_t* pk_lists;             // [2 parities][2: needs linearization, needs solve][B] trajectory indices
  unsigned int* pk_count;        // [2][2] list lengths
  unsigned long long* pk_mask;   // [B][C] spheres within reach of their hinge at the last evaluated states
  int32_t pk_mask_use, pk_pad_;  // the linearize kernel skips the spheres outside the masks (GPMP2B_PK_MASK=0: evaluates all)
};
// scalars of a trajectory in pk_state
enum PkScalar { PKS_LAMBDA = 0, PKS_ERROR, PKS_CURRENT_ERROR, PKS_LIN_COST_CHANGE, PKS_ITERATIONS, PKS_STATUS, PKS_SOLVED, PKS_COUNT = 16 };
__host__ __device__ inline int pk_even(int x) { return (x + 1) & ~1; }
__host__ __device__ inline int pk_state_size(int D, int N) { return 2 * pk_even(2 * D * N) + PKS_COUNT; }
__host__ __device__ inline int pk_row_stride(int D) { return pk_even(D * (D + 1) / 2 + D); }   // RS: doubles per configuration
__host__ __device__ inline size_t pk_mlist_size(int D, int N, int K) { return (size_t)pk_row_stride(D) * ((N - 1) * (K + 1) + 1); }
// shared-memory doubles of the streamed solve kernel: xs | g | dl | colbuf | ZT | ZB (the two coupling-block windows)
// | Hd -- the coupling blocks live in the M-list / the per-block global slab instead of shared memory
__host__ __device__ inline int pk_solve_smem(int D, int N) {
  const int b = 2 * D;
  return 3 * pk_even(N * b) + 144 + 2 * pk_even(b * b) + pk_even(N * (b * (b + 1) / 2));
}
// per resident solve block: the factored coupling blocks (transposed) kept for the back substitution
__host__ __device__ inline int pk_slab_size(int D, int N) { return pk_even((N - 1) * 4 * D * D) + 2; }
// shared-memory doubles per warp of the linearize (xs) and error (xs | dl) kernels
__host__ __device__ inline int pk_small_smem(int D, int N, bool with_dl) { return (with_dl ? 2 : 1) * pk_even(2 * D * N); }
// linearize kernel: xs + a 32-row staging buffer for the coalesced M-list stores
__host__ __device__ inline int pk_lin_smem(int D, int N) { return pk_even(2 * D * N) + 32 * pk_row_stride(D); }

// Pose2Vector states (optimizer_kernel_lie.cuh): the error kernel also holds the candidate states (xs | dl | cand);
// the linearize kernel holds xs | g | staging | Ho | Hd (the state-dependent GP-prior Hessian is assembled in place)
__host__ __device__ inline int lie_stage_per_config(int D);
__host__ __device__ inline int pk_lie_err_smem(int D, int N) { return 3 * pk_even(2 * D * N); }
__host__ __device__ inline int pk_lie_lin_smem(int D, int N) {
  const int b = 2 * D;
  return 2 * pk_even(N * b) + ((4 * lie_stage_per_config(D) + 32 + 1) & ~1) + (N - 1) * 24 + (N - 1) * b * b + pk_even(N * (b * (b + 1) / 2));
}

// H-path of the pipeline (pk_solve_mma.cuh): per-trajectory normal equations in HBM between the linearize and the solve
// kernel, [Ho: (N-1) b x b | Hd: N packed-lower | g: N b], in the pk_mlist buffer (whichever layout is larger sizes it)
__host__ __device__ inline size_t pk_hbuf_size(int D, int N) {
  const int b = 2 * D;
  return (size_t)pk_even((N - 1) * b * b + N * (b * (b + 1) / 2)) + pk_even(N * b);
}
// intervals per pass of the assembling linearize kernel (6 configurations each at obs_check_inter = 5: 30 lanes, the
// last state's unary factor rides in lane 30 of the last pass) and its shared-memory doubles per warp:
// xs | g | staging rows | carry | pass buffer (5 coupling blocks, 6 diagonal blocks)
#define PK_LINH_IPP 5
__host__ __device__ inline int pk_linh_smem(int D, int N) {
  const int b = 2 * D, T = D * (D + 1) / 2;
  return 2 * pk_even(N * b) + 32 * pk_row_stride(D) + pk_even(3 * T + 2 * D) + PK_LINH_IPP * b * b + pk_even((PK_LINH_IPP + 1) * (b * (b + 1) / 2));
}
// shared-memory doubles per trajectory of the tensor-core solve kernel (pk_solve_mma.cuh): g | dl | scratch | Ho | Hd
__host__ __device__ inline int pkm_smem_doubles(int D, int N) {
  const int b = 2 * D;
  return 2 * pk_even(N * b) + 64 + (N - 1) * b * b + pk_even(N * (b * (b + 1) / 2));
}

// ---- shared-memory layout of one trajectory (doubles); see optimizer_kernel.cuh ----
struct SmemLayout {
  int xs, g, dl, cand, Hd, Ho, stage, colbuf, geom, total;
};
// lie: Pose2Vector states (optimizer_kernel_lie.cuh) need a candidate-state array and a larger staging buffer
__host__ __device__ inline int lie_stage_per_config(int D) { return 4 * D * D + 36 + 4 + D; }
__host__ __device__ inline SmemLayout smem_layout(int D, int N, bool lie = false) {
  const int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b, T = D * (D + 1) / 2;
  SmemLayout L;
  int off = 0;
  auto even = [](int x) { return (x + 1) & ~1; };   // 16-byte alignment for double2 accesses
  L.xs = off; off += even(N * b);
  L.g = off; off += even(N * b);
  L.dl = off; off += even(N * b);
  L.cand = off; off += lie ? even(N * b) : 0;
  L.colbuf = off; off += 144;             // 2 (double buffer) x 2 (panels) x 36: column broadcast of the panel factorization
  L.stage = off; off += lie ? even(4 * lie_stage_per_config(D) + 32) : even(8 * (T + D));
  L.geom = off; off += lie ? (N - 1) * 24 : 0;   // LieOpt::geom: per-interval Logmap geometry
  L.Ho = off; off += (N - 1) * BB;        // Ho first: its blocks need 16-byte alignment; Hd follows contiguously
  L.Hd = off; off += even(N * BD);
  L.total = off;
  return L;
}
// size (doubles) of the H backup per resident warp
__host__ __device__ inline int h_backup_size(int D, int N) {
  const int b = 2 * D;
  return ((N * (b * (b + 1) / 2) + (N - 1) * b * b + 1) & ~1) + 2;   // even, 16-byte aligned slabs
}

enum { KMODE_OPTIMIZE = 0, KMODE_LINEARIZE = 1, KMODE_OBS_ERRORS = 2, KMODE_COLLISION_COST = 3 };

// kernel entry points are instantiated one translation unit per (robot kind, dof) -- kernels_inst.cu -- so that
// the build parallelises; this is the signature they all share and the lookup each unit exports
typedef void (*KernelFn)(const KRobot, const KSdf, const KSetting, const KProblem, const double*, int);
// opt: 0 Gauss-Newton, 1 LM, 2 Dogleg, -1 auxiliary kernel (linearize / obstacle-errors / collision-cost modes);
// + KOPT_GOAL: the instantiations that carry the optional workspace-goal / self-collision factors (vector-state robots only)
#define KOPT_GOAL 16
// phase-kernel pipeline of the LM optimizer (vector-state robots, default factors): linearize / assemble+solve / error+decision
#define KOPT_PK_LIN 32
#define KOPT_PK_SOLVE 33
#define KOPT_PK_ERR 34
#define KOPT_PK_SOLVE_MMA 35   // the solve phase on the FP64 tensor cores, two warps per trajectory (pk_solve_mma.cuh)
// ... with the assembly moved into the linearize kernel (obs_check_inter = 5): linearize + assemble -> H in HBM, load + solve
#define KOPT_PK_LINH 36
#define KOPT_PK_SOLVE_MMA_H 37
#define GPMP2B_DECLARE_LOOKUP(KIND, DD) KernelFn gpmp2b_lookup_##KIND##_##DD(int ndim, int opt);
