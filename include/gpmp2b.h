/*
 * gpmp2b.h -- C ABI of the B200-native GPMP2 batched trajectory-optimization hot path.
 *
 * This is the drop-in boundary.  Everything above it (the C++ facade in
 * include/gpmp2b/gpmp2.hpp, the ctypes mirror in gpmp2_b200/) only packs arguments; everything
 * below it is hand-written sm_100a CUDA.  Plain pointers and sizes only: no C++ types, no torch
 * types, no exceptions cross this boundary.
 *
 * Each entry point cites the reference interface (paths relative to the ori-drs/gpmp2 tree) it
 * replaces.  INTEGRATION.md shows the binding a gpmp2 maintainer would add.
 *
 * Flat layouts
 * ------------
 *  trajectory ("Values" of x(0..T), v(0..T)), per problem, T = total_step, N = T+1:
 *        [ x_0 .. x_T | v_0 .. v_T ],  each state D doubles, row-major, 2*N*D doubles total.
 *        This is the only flat layout the reference itself defines
 *        (gpmp2/utils/OpenRAVEutils.cpp:26-40, 50-66).  For Pose2Vector states
 *        x_i = (x, y, theta, q_1..q_n), D = 3+n.
 *  3-D SDF data: data[z][col][row]  (each z-slice is the reference's column-major Matrix,
 *        gpmp2/obstacle/SignedDistanceField.h:51,171; row = y index, col = x index).
 *  2-D SDF data: data[col][row]     (gpmp2/obstacle/PlanarSDF.h:41,118).
 */
#ifndef GPMP2B_H_
#define GPMP2B_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPMP2B_MAX_DOF 8      /* system dof D (arm: joints; mobile arm: 3 + joints) */
#define GPMP2B_MAX_SPHERES 64
#define GPMP2B_MAX_SELF_PAIRS 32

/* ---- status codes (return value of every call) ------------------------------------------- */
enum {
  GPMP2B_OK = 0,
  GPMP2B_ERR_INVALID_ARG = -1,   /* maps to std::runtime_error / std::invalid_argument in the facade */
  GPMP2B_ERR_CUDA = -2,
  GPMP2B_ERR_UNSUPPORTED = -3,
  GPMP2B_ERR_NO_DEVICE = -4
};

/* ---- per-problem status bits written to out_status ---------------------------------------- */
enum {
  GPMP2B_ST_CONVERGED_ABS = 1,   /* checkConvergence: absolute decrease <= 1e-5 [GTSAM default]   */
  GPMP2B_ST_CONVERGED_REL = 2,   /* checkConvergence: relative decrease <= rel_thresh             */
  GPMP2B_ST_MAX_ITER = 4,        /* iterations >= max_iter                                        */
  GPMP2B_ST_LAMBDA_MAXED = 8,    /* LM gave up: lambda >= lambdaUpperBound (1e5)                  */
  GPMP2B_ST_SOLVE_FAILED = 16,   /* non-SPD pivot (GTSAM IndeterminantLinearSystemException path) */
  GPMP2B_ST_ERROR_TOL = 32,      /* error <= errorTol (0) before/after iterating                  */
  GPMP2B_ST_ERR_INCREASED = 64   /* last step raised the error; last_values returned
                                    (gpmp2/planner/BatchTrajOptimizer.cpp:297-307)                */
};

/* robot kinds */
enum {
  GPMP2B_ROBOT_ARM = 0,
  GPMP2B_ROBOT_POSE2_MOBILE_ARM = 1,
  /* the other Pose2Vector robots of gpmp2/planner/BatchTrajOptimizer.cpp:92-128 */
  GPMP2B_ROBOT_POSE2_MOBILE_2ARMS = 2,       /* gpmp2/kinematics/Pose2Mobile2Arms.cpp:33-104        */
  GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_ARM = 3,  /* gpmp2/kinematics/Pose2MobileVetLinArm.cpp:31-94     */
  GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_2ARMS = 4 /* gpmp2/kinematics/Pose2MobileVetLin2Arms.cpp:34-114  */
};
/* optimizer kinds: same numbering as TrajOptimizerSetting::IterationType
 * (gpmp2/planner/TrajOptimizerSetting.h:21) */
enum { GPMP2B_OPT_GAUSS_NEWTON = 0, GPMP2B_OPT_LM = 1, GPMP2B_OPT_DOGLEG = 2 };
/* pointer-location flags for gpmp2b_batch_optimize & friends */
enum { GPMP2B_MEM_HOST = 0, GPMP2B_MEM_DEVICE = 1 };

/*
 * Robot description.  Replaces gpmp2::Arm (gpmp2/kinematics/Arm.h:47-55, DH a/alpha/d, base pose,
 * theta bias), gpmp2::BodySphere (gpmp2/kinematics/RobotModel.h:20-27), gpmp2::ArmModel
 * (gpmp2/kinematics/ArmModel.h:19) and gpmp2::Pose2MobileArmModel
 * (gpmp2/kinematics/Pose2MobileArm.h:41, Pose2MobileArmModel.h:19).
 */
typedef struct gpmp2b_robot_desc {
  int32_t kind;                 /* GPMP2B_ROBOT_*                                              */
  int32_t arm_dof;              /* number of DH joints                                         */
  int32_t n_spheres;
  int32_t reserved_;
  const double* a;              /* [arm_dof] DH a                                              */
  const double* alpha;          /* [arm_dof] DH alpha                                          */
  const double* d;              /* [arm_dof] DH d                                              */
  const double* theta_bias;     /* [arm_dof] or NULL (= zeros)                                 */
  double base_pose[16];         /* row-major 4x4.  ARM: Arm base_pose.  MOBILE: base_T_arm      */
  const int32_t* sphere_link;   /* [S] link id 0..nr_links-1 (mobile: 0 = vehicle base)        */
  const double* sphere_radius;  /* [S]                                                         */
  const double* sphere_center;  /* [S][3] centre in the link frame                             */
  /* ---- the two-arm / linear-actuator mobile manipulators (all zero for kinds 0 and 1) ----
   * State = Pose2Vector (x, y, theta | [z lift] | arm 1 joints | arm 2 joints); system dof = 3 + [1] + arm_dof + arm2_dof.
   * Links: 0 = vehicle, [1 = torso], then the joint frames of arm 1, then those of arm 2.
   * a / alpha / d / theta_bias hold the arm_dof joints of arm 1 followed by the arm2_dof joints of arm 2.
   *   MOBILE_2ARMS       base_pose = base_T_arm1,  base_pose2 = base_T_arm2
   *   MOBILE_VETLIN_ARM  base_pose = base_T_torso, base_pose2 = torso_T_arm;  torso = Trans(0, 0, +-z) * vehicle * base_T_torso
   *                      (liftBasePose3, gpmp2/kinematics/mobileBaseUtils.cpp:51-82; reverse_linact: -z)
   *   MOBILE_VETLIN_2ARMS base_pose = base_T_torso, base_pose2 = torso_T_arm1, base_pose3 = torso_T_arm2 */
  int32_t arm2_dof;
  int32_t reverse_linact;
  double base_pose2[16];
  double base_pose3[16];
} gpmp2b_robot_desc;

/*
 * Signed distance field.  Replaces gpmp2::SignedDistanceField(origin, cell_size, data)
 * (gpmp2/obstacle/SignedDistanceField.h:58-60) and gpmp2::PlanarSDF(origin, cell_size, data)
 * (gpmp2/obstacle/PlanarSDF.h:45-47).
 */
typedef struct gpmp2b_sdf_desc {
  int32_t ndim;                 /* 2 or 3                                                      */
  int32_t rows, cols, nz;       /* rows = y count, cols = x count, nz = z count (1 for 2-D)    */
  double origin[3];
  double cell_size;
  const double* data;           /* see "Flat layouts" above                                    */
} gpmp2b_sdf_desc;

/*
 * Mirrors gpmp2::TrajOptimizerSetting (gpmp2/planner/TrajOptimizerSetting.h:17-100, defaults
 * TrajOptimizerSetting.cpp:32-56) with the noise models flattened to what the setters can
 * produce (Isotropic sigma for the priors, Diagonal sigmas for the limits, full covariance Qc).
 */
typedef struct gpmp2b_setting {
  int32_t dof;
  int32_t total_step;
  double total_time;
  double conf_prior_sigma;
  double vel_prior_sigma;
  int32_t flag_pos_limit;
  int32_t flag_vel_limit;
  const double* joint_pos_limits_up;    /* [D] (NULL allowed if !flag_pos_limit)               */
  const double* joint_pos_limits_down;  /* [D]                                                 */
  const double* vel_limits;             /* [D] (NULL allowed if !flag_vel_limit)               */
  const double* pos_limit_thresh;       /* [D]                                                 */
  const double* vel_limit_thresh;       /* [D]                                                 */
  const double* pos_limit_sigma;        /* [D] Diagonal::Sigmas                                */
  const double* vel_limit_sigma;        /* [D]                                                 */
  double epsilon;
  double cost_sigma;
  int32_t obs_check_inter;
  int32_t opt_type;                     /* GPMP2B_OPT_*                                        */
  const double* Qc;                     /* [D*D] row-major covariance; NULL = identity         */
  int32_t opt_verbosity;                /* ignored on device; kept for struct parity           */
  int32_t final_iter_no_increase;       /* stored, never read -- like the reference
                                           (BatchTrajOptimizer-inl.h:83 always passes true)    */
  double rel_thresh;
  int32_t max_iter;
  int32_t reserved_;
  /* ---- optional workspace goal on the LAST support state (all zero = the reference's
   * BatchTrajOptimize graph).  One factor  e = position(joint frame goal_link)(x_T) - goal_pos,
   * Isotropic::Sigma(3, goal_sigma):  gpmp2::GoalFactorArm (gpmp2/kinematics/GoalFactorArm.h:47-77, goal_link =
   * arm_dof - 1) == gpmp2::GaussianPriorWorkspacePositionArm (GaussianPriorWorkspacePosition.h:46-76, any joint).
   * goal_keep_end_prior = 0: the factor REPLACES PriorFactor(x_T, end_conf), the hand-built graph of
   * matlab/Arm3GoalReachExample.m:104-108 (end_conf is then only the end of the straight-line initialisation);
   * 1: both factors.  The goal is shared by the B problems of a call (random restarts of one query).
   * goal_link < 0 = the last link frame.  Pose2MobileArm robots: GaussianPriorWorkspacePosition<Pose2MobileArmModel>,
   * link frames as Pose2MobileArm::forwardKinematics numbers them (0 = vehicle, 1..arm_dof = arm joint frames,
   * gpmp2/kinematics/Pose2MobileArm.cpp:30-108). */
  int32_t goal_enabled;
  int32_t goal_link;
  int32_t goal_keep_end_prior;
  int32_t reserved2_;
  double goal_sigma;
  double goal_pos[3];
  /* ---- optional self-collision factor on EVERY support state (0 = off):
   * gpmp2::SelfCollisionArm(x_i, arm, data) (gpmp2/obstacle/SelfCollision.h:38-128, SelfCollisionArm.h), data = n rows of
   * (sphere A id, sphere B id, epsilon, sigma): e_p = hinge(r_A + r_B + epsilon - |c_A - c_B|), Diagonal::Sigmas(sigma).
   * Sphere ids index gpmp2b_robot_desc's sphere arrays.  At most GPMP2B_MAX_SELF_PAIRS rows.  Arms and Pose2MobileArm. */
  int32_t n_self_collision;
  int32_t reserved3_;
  const double* self_collision_data;    /* [n_self_collision][4] row-major */
  /* ---- optional vehicle-dynamics factor on every support state of a Pose2MobileArm (0 = off):
   * gpmp2::VehicleDynamicsFactorPose2Vector(x_i, v_i, sigma) (gpmp2/dynamics/VehicleDynamicsFactorPose2Vector.h:46-79,
   * VehicleDynamics.h:19-28): e = v_i(1), the sideways (sliding) body velocity, Isotropic::Sigma(1, sigma) -- the graph
   * of matlab/MobileArm2FactorGraphExample.m:122-126.  GPMP2B_ERR_INVALID_ARG for arms. */
  double vehicle_dynamics_sigma;
  /* ---- optional workspace orientation prior on support states orient_state_first..orient_state_last (0 = off):
   * gpmp2::GaussianPriorWorkspaceOrientationArm(x_i, arm, orient_link, Rot3(orient_R), Isotropic::Sigma(3, orient_sigma))
   * (gpmp2/kinematics/GaussianPriorWorkspaceOrientation.h:40-72): e = Logmap(orient_R^T * R_link(x_i)) -- the
   * "keep the end effector upright" factors of matlab/WAMWorkspaceConstraintsExample.m:100-104.
   * orient_R row-major; orient_link < 0 = last link frame; link numbering as goal_link. */
  int32_t orient_enabled;
  int32_t orient_link;
  int32_t orient_state_first;
  int32_t orient_state_last;
  double orient_sigma;
  double orient_R[9];
  /* goal_enabled = 2: the workspace goal is a full pose, gpmp2::GaussianPriorWorkspacePoseArm(x_T, arm, goal_link,
   * Pose3(Rot3(goal_R), goal_pos), Isotropic::Sigma(6, goal_sigma)) (gpmp2/kinematics/GaussianPriorWorkspacePose.h:40-70):
   * e = Logmap(goal^-1 * T_link(x_T)) = [omega; u] -- the end-state factor of matlab/WAMWorkspaceConstraintsExample.m:94-96.
   * goal_R row-major; the other goal_* fields as for the position goal (goal_enabled = 1). */
  double goal_R[9];
  /* Per-problem workspace targets: the reference attaches these factors per graph (one query = one goal,
   * matlab/Arm3GoalReachExample.m:104-108, gpmp2/kinematics/GoalFactorArm.h:47-77), so a batch of DIFFERENT queries needs
   * one target per problem.  Each pointer is NULL (every problem uses the shared field above) or an array with one row
   * per problem of the call, in the call's memory space (`mem`: host or device pointers like start_conf):
   *   goal_pos_batch [B][3]  replaces goal_pos      (goal_enabled = 1 or 2)
   *   goal_R_batch   [B][9]  replaces goal_R        (goal_enabled = 2; row-major rotations, not re-validated per row)
   *   orient_R_batch [B][9]  replaces orient_R      (orient_enabled)
   * Ignored by the entry points that do not build the graph (gpmp2b_collision_cost, gpmp2b_obstacle_errors). */
  const double* goal_pos_batch;
  const double* goal_R_batch;
  const double* orient_R_batch;
  /* ---- replanning re-solve: the executed part of a plan is pinned and the rest re-optimized from the
   * previous solution, as gpmp2::ISAM2TrajOptimizer does between updates (gpmp2/planner/ISAM2TrajOptimizer-inl.h:121-194):
   *   fixConfigAndVel(k, conf, vel)   -> fix_enabled = 1, fix_state_index = k, fix_conf / fix_vel: PriorFactor(x_k, conf,
   *                                      conf_prior_model) + PriorFactor(v_k, vel, vel_prior_model) (-inl.h:160-168)
   *   changeGoalConfigAndVel(...)     -> new end_conf / end_vel rows of the call (-inl.h:121-141)
   *   initValues(previous result)     -> init_traj = the previous call's out_traj (warm start)
   * as ONE batched call over B independent replanning problems in lockstep (same k for all), with the optimizer of
   * opt_type run to max_iter -- a batch re-solve, not iSAM2's incremental Bayes-tree update.
   * fix_conf, fix_vel: [B][dof], one row per problem, in the call's memory space like start_conf. */
  int32_t fix_enabled;
  int32_t fix_state_index;
  const double* fix_conf;
  const double* fix_vel;
} gpmp2b_setting;

typedef struct gpmp2b_ctx gpmp2b_ctx;
typedef struct gpmp2b_robot gpmp2b_robot;   /* device-resident, immutable, owned by ctx */
typedef struct gpmp2b_sdf gpmp2b_sdf;       /* device-resident, immutable, owned by ctx */

/* ---- lifetime ----------------------------------------------------------------------------- */
int gpmp2b_create(int device, gpmp2b_ctx** out_ctx);
void gpmp2b_destroy(gpmp2b_ctx* ctx);
/* Last error message of this ctx (never NULL). */
const char* gpmp2b_last_error(const gpmp2b_ctx* ctx);
/* Library version string, for smoke checks on machines without a GPU. */
const char* gpmp2b_version(void);

/* ArmModel(arm, spheres) / Pose2MobileArmModel(marm, spheres): copy to device. */
int gpmp2b_robot_upload(gpmp2b_ctx* ctx, const gpmp2b_robot_desc* desc, gpmp2b_robot** out);
void gpmp2b_robot_free(gpmp2b_ctx* ctx, gpmp2b_robot* robot);
/* SignedDistanceField(...) / PlanarSDF(...): copy field to device (host pointer in desc). */
int gpmp2b_sdf_upload(gpmp2b_ctx* ctx, const gpmp2b_sdf_desc* desc, gpmp2b_sdf** out);
/*
 * Signed distance field from an occupancy grid on the device -- signedDistanceField3D / signedDistanceField2D
 * (matlab/+gpmp2/signedDistanceField3D.m:16-33, signedDistanceField2D.m): cells with occupancy > 0.75 are obstacles,
 * field = (EDT to the nearest obstacle cell - EDT to the nearest free cell) * cell_size with the exact Euclidean
 * distance transform (MATLAB bwdist), 1000 everywhere when there is no obstacle (or no free cell).
 * d->data holds the occupancy map as doubles in the SDF wire layout [z][col][row]; origin / cell_size as for an SDF.
 * single_precision != 0 reproduces MATLAB's arithmetic (bwdist returns single; the result is double(single));
 * 0 keeps everything in double (scipy's distance_transform_edt convention).
 * out (may be NULL): a device-resident handle like gpmp2b_sdf_upload's; out_field (may be NULL): the field on the host,
 * same layout as the input.
 */
int gpmp2b_sdf_from_occupancy(gpmp2b_ctx* ctx, const gpmp2b_sdf_desc* occupancy, int single_precision,
                              gpmp2b_sdf** out, double* out_field);
void gpmp2b_sdf_free(gpmp2b_ctx* ctx, gpmp2b_sdf* sdf);

/*
 * The hot path.  Replaces, for B independent problems at once,
 *   gpmp2::BatchTrajOptimize2DArm            (gpmp2/planner/BatchTrajOptimizer.h:43-47,  .cpp:40-50)
 *   gpmp2::BatchTrajOptimize3DArm            (BatchTrajOptimizer.h:62-66,  .cpp:53-63)
 *   gpmp2::BatchTrajOptimizePose2MobileArm2D (BatchTrajOptimizer.h:81-85,  .cpp:66-76)
 *   gpmp2::BatchTrajOptimizePose2MobileArm   (BatchTrajOptimizer.h:100-104, .cpp:79-89)
 * i.e. internal::BatchTrajOptimize (BatchTrajOptimizer-inl.h:19-84) + gpmp2::optimize
 * (BatchTrajOptimizer.cpp:212-308) + the GTSAM optimizer underneath.  Which of the four is
 * selected by (robot kind, sdf ndim).
 *
 *  start_conf,end_conf [B][D]; start_vel,end_vel [B][D]; init_traj,out_traj [B][2*N*D].
 *  init_traj may be NULL: the straight line from start_conf to end_conf (initArmTrajStraightLine /
 *  initPose2VectorTrajStraightLine, TrajUtils.cpp:23-73) is then built on the device -- no trajectory upload.
 *  out_error [B] final graph error (0.5 * sum whitened^2); out_coll_cost [B] = CollisionCost*()
 *  of the result (BatchTrajOptimizer-inl.h:87-100); out_iters [B]; out_status [B] bit mask.
 *  Any of the four out_* scalars arrays may be NULL.  mem = GPMP2B_MEM_HOST: all pointers are host
 *  buffers (copies are part of the call); GPMP2B_MEM_DEVICE: all are device pointers, the call is
 *  stream-ordered on `cuda_stream` (a cudaStream_t, NULL = default stream) and asynchronous.
 *  Calls on one ctx share its device scratch (H template, H-backup slabs, work queues): a call issued on a
 *  different stream than the previous one first waits, on the device, for the previous call's kernels (an event
 *  recorded at the end of every call), so calls on one ctx never overlap each other; use one ctx per stream
 *  to run calls concurrently.  A ctx must not be used from two host threads at once.
 */
int gpmp2b_batch_optimize(gpmp2b_ctx* ctx, const gpmp2b_robot* robot, const gpmp2b_sdf* sdf,
                          const gpmp2b_setting* setting, int64_t B,
                          const double* start_conf, const double* start_vel,
                          const double* end_conf, const double* end_vel,
                          const double* init_traj, double* out_traj,
                          double* out_error, double* out_coll_cost,
                          int32_t* out_iters, int32_t* out_status,
                          int mem, void* cuda_stream);

/*
 * The same call spread over several GPUs of one box (the problems of a batch are independent, so the
 * batch is cut into n_dev contiguous shards with no data-path collective).  ctxs[i] / robots[i] / sdfs[i] are one
 * context per device (gpmp2b_create(device_i, ..)) with the robot and the field uploaded to each; shard i -- problems
 * [i B / n_dev, (i + 1) B / n_dev) -- runs on ctxs[i], driven by its own host thread.
 *   mem = GPMP2B_MEM_HOST  : all pointers are host buffers; every device copies its shard straight from / to the
 *                            caller's arrays (chunk-pipelined, as in gpmp2b_batch_optimize), nothing passes through GPU 0.
 *   mem = GPMP2B_MEM_DEVICE: all pointers are buffers on ctxs[0]'s device.  The other devices pull their shard of the
 *                            inputs from it and push their results back into the caller's output arrays with peer copies
 *                            (NVLink / NVSwitch where the box has it) -- the gather of results and costs to one GPU that
 *                            BASELINE.json's north_star names, without NCCL or a second process.  The call returns after
 *                            all shards have landed.
 * Two contexts on the same device are allowed (they then simply share it).  Returns the first shard's error, if any;
 * gpmp2b_last_error(ctxs[i]) has the message of shard i.
 * Replaces a loop of gpmp2::BatchTrajOptimize* calls over the queries of a batch (gpmp2/planner/BatchTrajOptimizer.h:43-104).
 */
int gpmp2b_batch_optimize_multi(int n_dev, gpmp2b_ctx* const* ctxs, const gpmp2b_robot* const* robots,
                                const gpmp2b_sdf* const* sdfs, const gpmp2b_setting* setting, int64_t B,
                                const double* start_conf, const double* start_vel,
                                const double* end_conf, const double* end_vel,
                                const double* init_traj, double* out_traj,
                                double* out_error, double* out_coll_cost,
                                int32_t* out_iters, int32_t* out_status, int mem);

/*
 * CollisionCost2DArm / 3DArm / Pose2MobileArm2D / Pose2MobileArm
 * (gpmp2/planner/BatchTrajOptimizer.h:135-185, -inl.h:87-100): epsilon = 0, unwhitened sum over
 * the support states of the unary obstacle factor error.
 */
int gpmp2b_collision_cost(gpmp2b_ctx* ctx, const gpmp2b_robot* robot, const gpmp2b_sdf* sdf,
                          const gpmp2b_setting* setting, int64_t B, const double* traj,
                          double* out_cost, int mem, void* cuda_stream);

/*
 * Parity/debug entry: one linearization of the whole factor graph at `traj`
 * (what NonlinearFactorGraph::linearize + error produce for the graph of
 * BatchTrajOptimizer-inl.h:36-81), returned in block-tridiagonal form, b = 2*D:
 *  out_Hdiag [B][N][b][b]   full symmetric diagonal blocks of H = J^T J (undamped)
 *  out_Hoff  [B][N-1][b][b] H_{i,i+1} (rows: state i, cols: state i+1)
 *  out_g     [B][N][b]      gradient J^T W e, per state ordered [x_i ; v_i]
 *  out_error [B]            0.5 * sum of squared whitened errors
 * Any output may be NULL.
 */
int gpmp2b_linearize(gpmp2b_ctx* ctx, const gpmp2b_robot* robot, const gpmp2b_sdf* sdf,
                     const gpmp2b_setting* setting, int64_t B,
                     const double* start_conf, const double* start_vel,
                     const double* end_conf, const double* end_vel, const double* traj,
                     double* out_Hdiag, double* out_Hoff, double* out_g, double* out_error,
                     int mem, void* cuda_stream);

/*
 * Parity/debug entry: unwhitened obstacle-factor errors of every collision-checked configuration,
 * i.e. ObstacleSDFFactor::evaluateError (gpmp2/obstacle/ObstacleSDFFactor-inl.h:18-55) at each
 * support state and ObstacleSDFFactorGP::evaluateError (ObstacleSDFFactorGP-inl.h:18-75) at each
 * interpolated state, plus the whitened Jacobian row products are checked through gpmp2b_linearize.
 *  out_err [B][C][S], C = N + (N-1)*obs_check_inter, configurations ordered
 *  (i=0,j=0), (i=0,j=1..K), (i=1,j=0), ... , (i=N-1,j=0)   with j=0 the support state.
 *  out_centers [B][C][S][3] sphere centres (RobotModel::sphereCenters, RobotModel-inl.h:12-40) or NULL.
 */
int gpmp2b_obstacle_errors(gpmp2b_ctx* ctx, const gpmp2b_robot* robot, const gpmp2b_sdf* sdf,
                           const gpmp2b_setting* setting, int64_t B, const double* traj,
                           double* out_err, double* out_centers, int mem, void* cuda_stream);

/*
 * Trajectory utilities either side of the planner (gpmp2/planner/TrajUtils.cpp), batched on the device.
 *
 * gpmp2b_init_straight_line: initArmTrajStraightLine (TrajUtils.cpp:23-48) for robot_kind = GPMP2B_ROBOT_ARM,
 *   initPose2VectorTrajStraightLine (:51-73) for GPMP2B_ROBOT_POSE2_MOBILE_ARM (start/end = (x, y, theta, q...),
 *   pose part interpolate<Pose2>(start, end, i / total_step)); v_i = (end - start) / total_step.
 *   start_conf, end_conf: [B][dof];  out_traj: [B][2 * (total_step + 1) * dof] in the wire layout.
 *
 * gpmp2b_interpolate_traj: interpolateArmTraj (TrajUtils.cpp:158-196) / interpolatePose2MobileArmTraj (:199-237),
 *   i.e. GaussianProcessInterpolator{Linear,Pose2Vector}::interpolatePose + interpolateVelocity at inter_step points
 *   per interval of states start_index .. end_index (the 4-argument interpolateArmTraj overload, :96-155, is
 *   start_index = 0, end_index = total_step).  Qc ([dof][dof] or NULL = identity) must be invertible; the result
 *   does not depend on it otherwise.  out_traj: [B][2 * Nout * dof], Nout = (end_index - start_index) * (inter_step + 1) + 1.
 *
 * gpmp2b_select_best: best of R restarts for each of G queries (problems g*R .. g*R + R - 1 of a batch): the index
 *   of the smallest final error among restarts whose collision cost is <= coll_tol (coll_cost may be NULL = all
 *   admissible); if none is admissible, the smallest error overall with out_feasible[g] = 0.  out_best: [G] indices
 *   into the batch (-1 if every error is NaN); out_feasible: [G] or NULL.
 */
int gpmp2b_init_straight_line(gpmp2b_ctx* ctx, int robot_kind, int dof, int total_step, int64_t B,
                              const double* start_conf, const double* end_conf, double* out_traj,
                              int mem, void* cuda_stream);
int gpmp2b_interpolate_traj(gpmp2b_ctx* ctx, int robot_kind, int dof, int total_step, double delta_t,
                            const double* Qc, int inter_step, int start_index, int end_index, int64_t B,
                            const double* traj, double* out_traj, int mem, void* cuda_stream);
int gpmp2b_select_best(gpmp2b_ctx* ctx, int64_t G, int64_t R, const double* error, const double* coll_cost,
                       double coll_tol, int64_t* out_best, int32_t* out_feasible, int mem, void* cuda_stream);

/*
 * Measured device peaks for the roofline (bench.py): a dependent-free DFMA loop and an
 * L2-resident random 32-byte-sector gather.  Results in out[0] = FP64 TFLOP/s (FMA = 2 flops),
 * out[1] = L2 gather GB/s with 8-byte loads (useful bytes), out[2] = L2 gather GB/s with one 256-bit load per lane
 * (32-byte quad cells, the SDF access pattern).
 */
int gpmp2b_measure_peaks(gpmp2b_ctx* ctx, double* out3);

/* Kernel-launch counter (bench.py "gpu_launches"): number of kernels this ctx launched so far. */
int64_t gpmp2b_launch_count(const gpmp2b_ctx* ctx);
/* Device time (ms, CUDA events on the launching stream) of the optimizer kernel of the most
 * recent gpmp2b_batch_optimize call on this ctx, and the LM/GN iterations it executed summed
 * over the batch (linearize count).  Valid after the stream has been synchronized. */
int gpmp2b_last_kernel_stats(gpmp2b_ctx* ctx, double* out_kernel_ms, int64_t* out_linearizations,
                             int64_t* out_solves, int64_t* out_error_evals);

#ifdef __cplusplus
}
#endif
#endif /* GPMP2B_H_ */
