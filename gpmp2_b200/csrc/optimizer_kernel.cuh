// optimizer_kernel.cuh -- the fused batched trajectory optimizer: ONE WARP PER TRAJECTORY.
//
// Replaces, per trajectory, the whole of internal::BatchTrajOptimize (gpmp2/planner/
// BatchTrajOptimizer-inl.h:19-84: graph of priors + unary obstacle + GP obstacle + GP prior [+ limit]
// factors), gpmp2::optimize (gpmp2/planner/BatchTrajOptimizer.cpp:212-308) and the GTSAM optimizer
// underneath (linearize -> damped normal equations -> solve -> retract -> error, LM lambda logic,
// checkConvergence).  Vector-valued states (ArmModel robots); the Pose2Vector variant lives in
// optimizer_kernel_lie.cuh.
//
// Data layout (shared memory, per warp, doubles; b = 2D, BD = b(b+1)/2, BB = b*b):
//   xs [N*b]   current states, state-major  s_i = [x_i ; v_i]
//   g  [N*b]   gradient J^T W e
//   dl [N*b]   rhs / forward-substituted rhs / delta
//   Hd [N*BD]  diagonal blocks of H, packed lower;   after factorization: L_ii (diagonal stored as 1/l_kk)
//   Ho [(N-1)*BB]  H_{i,i+1} row-major (rows: state i, cols: state i+1); after factorization the
//                  coupling factor of the two-sided sweep, row-major (see solve())
//   stage [8*(T+D)]  staging of per-configuration (M, c) between the configuration-parallel and the
//                    entry-parallel phase;  colbuf [128]
// H never leaves the SM except as a backup copy (L2-resident, per resident warp) used to restore it
// when LM rejects a step and retries with a larger lambda.
#pragma once
#include <math_constants.h>
#include "device_model.cuh"

#ifndef GPMP2B_BACKSUB_REG
#define GPMP2B_BACKSUB_REG 1
#endif
#ifndef GPMP2B_ALIGNED_ACC
#define GPMP2B_ALIGNED_ACC 1
#endif
#ifndef GPMP2B_SCHUR_REMAP
#define GPMP2B_SCHUR_REMAP 1
#endif
#ifndef GPMP2B_SCHUR_ROTATE
#define GPMP2B_SCHUR_ROTATE 0
#endif
#ifndef GPMP2B_RSQRT_HALLEY
#define GPMP2B_RSQRT_HALLEY 1
#endif
// EXTRA: the instantiation carries the optional factors of hand-built graphs -- workspace goal (goal_pass) and
// self-collision (self_pass), each still switched by the setting at run time.  A compile-time switch: with only a
// run-time test the out-of-line call cost the default WAM kernel 27 % (registers saved around the call site).
// experiment switch: which optional factors the EXTRA variants compile in (bit 0 position goal, 1 self-collision,
// 2 orientation prior, 3 pose goal)
#ifndef GPMP2B_LIN_ASYNC
#define GPMP2B_LIN_ASYNC 0
#endif
#ifndef GPMP2B_EXTRA_MASK
#define GPMP2B_EXTRA_MASK 15
#endif
#ifndef GPMP2B_EXTRA_NOINLINE
#define GPMP2B_EXTRA_NOINLINE 0
#endif
#if GPMP2B_EXTRA_NOINLINE
#define GPMP2B_EXTRA_FN __noinline__
#else
#define GPMP2B_EXTRA_FN __forceinline__
#endif
template <int D, int NDIM, bool EXTRA = false>
struct VecOpt {
  static constexpr int b = 2 * D;
  static constexpr int BD = b * (b + 1) / 2;
  static constexpr int BB = b * b;
  static constexpr int T = D * (D + 1) / 2;
  static constexpr int STG = T + D;
  static constexpr int Dim = D;
  static constexpr int NDim = NDIM;
  static constexpr bool LIE = false;

  const KRobot& rb;
  const KSdf& sdf;
  const KSetting& st;
  const double* __restrict__ hconst;
  int lane, N, K, C;
  double *xs, *g, *dl, *Hd, *Ho, *stage, *colbuf;
  double *ZT = nullptr, *ZB = nullptr;   // streamed solve (layout 3): windows of the top / bottom sweep's coupling block
  const double *start_conf, *start_vel, *end_conf, *end_vel;   // this problem's
  // workspace targets of this problem: the setting's shared values, or this problem's rows of gpmp2b_setting.*_batch
  const double *goal_pos_p = nullptr, *goal_R_p = nullptr, *orient_R_p = nullptr;
  const double *fix_conf = nullptr, *fix_vel = nullptr;   // this problem's fixed-state targets (st.fix_enabled)
  int tp, tq;   // lane's (p, q) of packed entry m = lane (p >= q), valid if lane < T
  int err_scratch_off = 0;    // doubles at the start of the H storage that the error pass must not use as scratch (Dogleg: dx_n)
  bool no_err_scratch = false;   // no H storage to borrow (error kernel of the phase pipeline): gather through registers
  unsigned long long* mask_out = nullptr;   // eval_error<CAND, true>: this trajectory's per-configuration sphere masks (global)
  double* escr = nullptr;        // ... unless the kernel brings its own landing zone for the asynchronous gathers:
  int escr_chunk = 0;            //     escr_chunk spheres x 3 KB (layout 2 with extra shared memory behind xs | dl)
  int sch_r, sch_c0, sch_n;   // Schur update: this lane owns entries (sch_r, sch_c0 .. sch_c0 + sch_n - 1) of a packed block
  int sch_rot = 0;            // first entry of the run this lane visits (GPMP2B_SCHUR_ROTATE)
#ifdef GPMP2B_PHASE_TIMING
  long long pt_cfg = 0, pt_acc = 0, pt_init = 0;
#endif

  // layout: 0 = the full per-trajectory layout (smem_layout); phase-kernel pipeline: 1 = xs only (linearize kernel),
  // 2 = xs | dl (error kernel) -- the other arrays do not exist there and the error pass gathers through registers
  __device__ VecOpt(const KRobot& rb_, const KSdf& sdf_, const KSetting& st_, const double* hc, double* smem, bool lie = false,
                    int layout = 0)
      : rb(rb_), sdf(sdf_), st(st_), hconst(hc) {
    lane = threadIdx.x & 31;
    N = st.N; K = st.K;
    C = (N - 1) * (K + 1) + 1;
    if (layout == 0) {
      const SmemLayout L = smem_layout(D, N, lie);
      xs = smem + L.xs; g = smem + L.g; dl = smem + L.dl; Hd = smem + L.Hd; Ho = smem + L.Ho;
      stage = smem + L.stage; colbuf = smem + L.colbuf;
    } else if (layout == 3) {   // streamed solve kernel (pk_solve_smem): no Ho, no staging buffer
      int off = 0;
      xs = smem + off; off += pk_even(N * b);
      g = smem + off; off += pk_even(N * b);
      dl = smem + off; off += pk_even(N * b);
      colbuf = smem + off; off += 144;
      ZT = smem + off; off += pk_even(BB);
      ZB = smem + off; off += pk_even(BB);
      Hd = smem + off;
      Ho = stage = nullptr;
      no_err_scratch = true;
    } else {
      xs = smem; dl = smem + pk_even(N * b);
      g = Hd = Ho = stage = colbuf = nullptr;
      no_err_scratch = true;
    }
    // Schur updates: row r of a packed-lower block is split into ceil((r + 1) / 4) runs of <= 4 entries, one run per
    // lane (b = 14: 4*1 + 4*2 + 4*3 + 2*4 = exactly 32 runs), so a lane loads its row of Z once for all its entries
    {
      sch_r = 0; sch_c0 = 0; sch_n = 0;
      int first = 0;
      for (int r = 0; r < b; r++) {
        const int runs = (r + 4) / 4;
        if (lane >= first && lane < first + runs) { sch_r = r; sch_c0 = 4 * (lane - first); sch_n = min(4, r + 1 - sch_c0); }
        first += runs;
      }
      sch_rot = (sch_c0 >= 8 && sch_n > 1) ? (sch_n == 4 ? 2 : 1) : 0;
#if GPMP2B_SCHUR_REMAP
      // b = 14: rows of Z that are 8 apart share their four banks (112-byte rows).  Same 32 runs, dealt to the lanes so
      // that within a quarter-warp (one LDS.128 wavefront) neither the lanes' own rows nor the rows c0 + j they walk are
      // ever 8 apart:  Q0 = (r 0..7, c0 0);  Q1 = (r 8..11, c0 0), (r 4..7, c0 4);  Q2 = (r 12..13, c0 0), (r 8..13, c0 4);
      // Q3 = (r 8..13, c0 8), (r 12..13, c0 12).  No extra instruction in the update itself.
      if (b == 14) {
        if (lane < 8) { sch_r = lane; sch_c0 = 0; }
        else if (lane < 12) { sch_r = lane; sch_c0 = 0; }
        else if (lane < 16) { sch_r = lane - 8; sch_c0 = 4; }
        else if (lane < 18) { sch_r = lane - 4; sch_c0 = 0; }
        else if (lane < 24) { sch_r = lane - 10; sch_c0 = 4; }
        else if (lane < 30) { sch_r = lane - 16; sch_c0 = 8; }
        else { sch_r = lane - 18; sch_c0 = 12; }
        sch_n = min(4, sch_r + 1 - sch_c0);
        sch_rot = 0;
      }
#endif
    }
    // (p, q), p >= q, of packed entry m = lane (closed form so that rematerialising it is cheap)
    tp = (int)((sqrtf(8.0f * (float)lane + 1.0f) - 1.0f) * 0.5f);
    if (tp * (tp + 1) / 2 > lane) tp--;
    if ((tp + 1) * (tp + 2) / 2 <= lane) tp++;
    tq = lane - tp * (tp + 1) / 2;
  }

  template <bool CAND>
  __device__ __forceinline__ double sv(int idx) const { return CAND ? xs[idx] + dl[idx] : xs[idx]; }

  // configuration (i, j): support state (j == 0) or GP-interpolated state
  // (GaussianProcessInterpolatorLinear::interpolatePose, gpmp2/gp/GaussianProcessInterpolatorLinear.h:62-84,
  //  with every D x D block of Lambda/Psi a scalar multiple of I -- SURVEY.md 7.1-3).  Coordinate d on demand.
  template <bool CAND>
  struct QFun {
    const double *xs, *dl;
    int o0, o1;            // offsets of s_i and s_{i+1} (clamped) in xs
    double w0, w1, w2, w3;
    __device__ __forceinline__ double at(int idx) const { return CAND ? xs[idx] + dl[idx] : xs[idx]; }
    __device__ __forceinline__ double operator()(int d) const {
      double v = w0 * at(o0 + d);
      v = fma(w1, at(o0 + D + d), v);
      v = fma(w2, at(o1 + d), v);
      v = fma(w3, at(o1 + D + d), v);
      return v;
    }
  };
  template <bool CAND>
  __device__ __forceinline__ QFun<CAND> config_state(int i, int j) const {
    QFun<CAND> f;
    f.xs = xs; f.dl = dl;
    f.o0 = i * b;
    f.o1 = min(i + 1, N - 1) * b;
    const int jj = max(j - 1, 0);
    f.w0 = j ? st.gpw[jj][0] : 1.0; f.w1 = j ? st.gpw[jj][1] : 0.0;
    f.w2 = j ? st.gpw[jj][2] : 0.0; f.w3 = j ? st.gpw[jj][3] : 0.0;
    return f;
  }

  // ---- per-(state, dof) pass: priors (PriorFactor), GP prior (GaussianProcessPriorLinear.h:57-83),
  //      limit hinges (JointLimitFactorVector.h:62-79, VelocityLimitFactorVector.h:62-79).
  //      GRAD: write g (overwrites) and add limit curvature to Hd.  Returns this lane's error share. ----
  template <bool CAND, bool GRAD>
  __device__ __forceinline__ double state_pass() {
    double eacc = 0.0;
    const double dt = st.delta_t;
    const double q11 = st.qi[0][0], q12 = st.qi[0][1], q22 = st.qi[1][1];
#pragma unroll 1
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      double gx = 0.0, gv = 0.0;
      // interval (i, i+1): e = Phi s_i - s_{i+1}; u = Q^-1 e; g_i += Phi^T u; error += 0.5 e.u
      if (i < N - 1) {
        const double exd = (sv<CAND>(i * b + d) + dt * sv<CAND>(i * b + D + d)) - sv<CAND>((i + 1) * b + d);
        const double evd = sv<CAND>(i * b + D + d) - sv<CAND>((i + 1) * b + D + d);
        double ux, uv;
        if (st.qc_identity) {   // Qc = I (the library default): Q^-1 e is two scalar combinations
          ux = fma(q11, exd, q12 * evd);
          uv = fma(q12, exd, q22 * evd);
        } else {
          ux = 0.0; uv = 0.0;
#pragma unroll 1
          for (int k = 0; k < D; k++) {
            const double qc = st.Qc_inv[d * D + k];
            const double ex = (sv<CAND>(i * b + k) + dt * sv<CAND>(i * b + D + k)) - sv<CAND>((i + 1) * b + k);
            const double ev = sv<CAND>(i * b + D + k) - sv<CAND>((i + 1) * b + D + k);
            ux = fma(qc, fma(q11, ex, q12 * ev), ux);
            uv = fma(qc, fma(q12, ex, q22 * ev), uv);
          }
        }
        eacc += 0.5 * fma(exd, ux, evd * uv);
        gx += ux;
        gv += fma(dt, ux, uv);
      }
      if (GRAD && i > 0) {   // interval (i-1, i): g_i -= u
        double ux, uv;
        if (st.qc_identity) {
          const double ex = (sv<CAND>((i - 1) * b + d) + dt * sv<CAND>((i - 1) * b + D + d)) - sv<CAND>(i * b + d);
          const double ev = sv<CAND>((i - 1) * b + D + d) - sv<CAND>(i * b + D + d);
          ux = fma(q11, ex, q12 * ev);
          uv = fma(q12, ex, q22 * ev);
        } else {
          ux = 0.0; uv = 0.0;
#pragma unroll 1
          for (int k = 0; k < D; k++) {
            const double qc = st.Qc_inv[d * D + k];
            const double ex = (sv<CAND>((i - 1) * b + k) + dt * sv<CAND>((i - 1) * b + D + k)) - sv<CAND>(i * b + k);
            const double ev = sv<CAND>((i - 1) * b + D + k) - sv<CAND>(i * b + D + k);
            ux = fma(qc, fma(q11, ex, q12 * ev), ux);
            uv = fma(qc, fma(q12, ex, q22 * ev), uv);
          }
        }
        gx -= ux;
        gv -= uv;
      }
      if (i == 0 || i == N - 1) {   // PriorFactor on x_i, v_i (BatchTrajOptimizer-inl.h:41-48)
        const double pc = (i == 0 ? start_conf : end_conf)[d], pv = (i == 0 ? start_vel : end_vel)[d];
        const double dx = sv<CAND>(i * b + d) - pc, dv = sv<CAND>(i * b + D + d) - pv;
        const double cw = i == 0 ? st.conf_prior_w : st.end_conf_prior_w;   // 0 when a workspace goal replaces the prior
        eacc += 0.5 * (cw * dx * dx + st.vel_prior_w * dv * dv);
        gx = fma(cw, dx, gx);
        gv = fma(st.vel_prior_w, dv, gv);
      }
      if (st.flag_pos_limit) {
        const double p = sv<CAND>(i * b + d), lo = st.pos_lo[d] + st.pos_th[d], hi = st.pos_hi[d] - st.pos_th[d];
        double e = 0.0, h = 0.0;
        if (p < lo) { e = lo - p; h = -1.0; }
        else if (p <= hi) { e = 0.0; h = 0.0; }
        else { e = p - hi; h = 1.0; }
        eacc += 0.5 * st.pos_w[d] * e * e;
        if (GRAD) {
          gx = fma(st.pos_w[d] * h, e, gx);
          Hd[i * BD + d * (d + 1) / 2 + d] += st.pos_w[d] * h * h;
        }
      }
      if (st.flag_vel_limit) {
        const double p = sv<CAND>(i * b + D + d), lo = -st.vel_lim[d] + st.vel_th[d], hi = st.vel_lim[d] - st.vel_th[d];
        double e = 0.0, h = 0.0;
        if (p < lo) { e = lo - p; h = -1.0; }
        else if (p <= hi) { e = 0.0; h = 0.0; }
        else { e = p - hi; h = 1.0; }
        eacc += 0.5 * st.vel_w[d] * e * e;
        if (GRAD) {
          gv = fma(st.vel_w[d] * h, e, gv);
          const int r = D + d;
          Hd[i * BD + r * (r + 1) / 2 + r] += st.vel_w[d] * h * h;
        }
      }
      if (GRAD) { g[i * b + d] = gx; g[i * b + D + d] = gv; }
    }
    return eacc;
  }

  // ---- frame walker shared by the optional-factor passes: base frame of the chain [X Y Z | o] for state reader sf(k).
  //      KIND 0: the arm's base pose.  KIND 1 (Pose2MobileArm): the vehicle frame Rz(theta), (x, y, 0) first -- veh()
  //      is called on it (spheres of link 0, pseudo-joint lines) -- then vehicle * base_T_arm. ----
  template <int KIND, class SF, class VF>
  __device__ __forceinline__ void chain_base(const SF& sf, double (&X)[3], double (&Y)[3], double (&Z)[3], double (&o)[3], VF&& veh) const {
    if (KIND == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        X[k] = rb.base[k * 4 + 0]; Y[k] = rb.base[k * 4 + 1]; Z[k] = rb.base[k * 4 + 2]; o[k] = rb.base[k * 4 + 3];
      }
    } else {
      double sn, cs;
      fast_sincos(sf(2), sn, cs);
      X[0] = cs; X[1] = sn; X[2] = 0.0;
      Y[0] = -sn; Y[1] = cs; Y[2] = 0.0;
      Z[0] = 0.0; Z[1] = 0.0; Z[2] = 1.0;
      o[0] = sf(0); o[1] = sf(1); o[2] = 0.0;
      veh();
      double nX[3], nY[3], nZ[3], no[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        nX[k] = X[k] * rb.base[0] + Y[k] * rb.base[4] + Z[k] * rb.base[8];
        nY[k] = X[k] * rb.base[1] + Y[k] * rb.base[5] + Z[k] * rb.base[9];
        nZ[k] = X[k] * rb.base[2] + Y[k] * rb.base[6] + Z[k] * rb.base[10];
        no[k] = fma(Z[k], rb.base[11], fma(Y[k], rb.base[7], fma(X[k], rb.base[3], o[k])));
      }
#pragma unroll
      for (int k = 0; k < 3; k++) { X[k] = nX[k]; Y[k] = nY[k]; Z[k] = nZ[k]; o[k] = no[k]; }
    }
  }
  // one DH step: T_{j+1} = T_j Rz(q + bias_j) Trans(a_j, 0, d_j) Rx(alpha_j)  (Arm.cpp:24-27, Arm.h:93-98)
  __device__ __forceinline__ void chain_step(int j, double q, double (&X)[3], double (&Y)[3], double (&Z)[3], double (&o)[3]) const {
    double sn, cs;
    fast_sincos(q + rb.bias[j], sn, cs);
    const double ca = rb.ca[j], sa = rb.sa[j], aj = rb.a[j], dj = rb.d[j];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double xn = fma(cs, X[k], sn * Y[k]);
      const double yn = fma(cs, Y[k], -sn * X[k]);
      o[k] = fma(dj, Z[k], fma(aj, xn, o[k]));
      const double y2 = fma(ca, yn, sa * Z[k]);
      const double z2 = fma(ca, Z[k], -sa * yn);
      X[k] = xn; Y[k] = y2; Z[k] = z2;
    }
  }

  // ---- optional workspace goal on the last support state (gpmp2b_setting.goal_*):
  //      GoalFactorArm::evaluateError (kinematics/GoalFactorArm.h:52-70) == GaussianPriorWorkspacePosition::evaluateError
  //      (GaussianPriorWorkspacePosition.h:54-71): e = origin of link frame goal_link - goal, Isotropic sigma.
  //      Every lane walks the same chain (warp-uniform; one factor per trajectory, off the hot loops); lane k keeps the
  //      line of (pseudo-)joint k, so column k of the Jacobian is z_k x (p - o_k) -- the reference's R * (T^-1 dT/dq_k)^v
  //      chain in world coordinates -- or the body axis for the two prismatic directions of the Pose2 chart.
  //      sf(k): coordinate k of x_T.  GRAD: add J^T J / sigma^2 to the position block of Hd[N-1] and J^T e / sigma^2 to g
  //      (after the per-state pass wrote g).  Returns the error share (lane 0 only). ----
  template <int KIND, bool GRAD, class SF>
  __device__ GPMP2B_EXTRA_FN double goal_eval(const SF& sf) {
    constexpr int NB = (KIND == 1) ? 3 : 0;
    double X[3], Y[3], Z[3], o[3], zk[3] = {0.0, 0.0, 0.0}, ok[3] = {0.0, 0.0, 0.0}, pk[3] = {0.0, 0.0, 0.0};
    chain_base<KIND>(sf, X, Y, Z, o, [&]() {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (lane == 0) pk[k] = X[k];
        if (lane == 1) pk[k] = Y[k];
        if (lane == 2) { zk[k] = Z[k]; ok[k] = o[k]; }
      }
    });
    const int narm = (KIND == 1) ? st.goal_link : st.goal_link + 1;   // arm joints up to the constrained frame
    double p[3] = {o[0], o[1], o[2]};
    if (KIND == 1 && narm == 0) { p[0] = sf(0); p[1] = sf(1); p[2] = 0.0; }   // link 0 = the vehicle frame itself
#pragma unroll 1
    for (int j = 0; j < narm; j++) {
      if (lane == NB + j) {
#pragma unroll
        for (int k = 0; k < 3; k++) { zk[k] = Z[k]; ok[k] = o[k]; }
      }
      chain_step(j, sf(NB + j), X, Y, Z, o);
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = o[k];
    }
    const double e0 = p[0] - goal_pos_p[0], e1 = p[1] - goal_pos_p[1], e2 = p[2] - goal_pos_p[2];
    if (GRAD) {
      const int last = (N - 1) * b;
      const double rx = p[0] - ok[0], ry = p[1] - ok[1], rz = p[2] - ok[2];
      const bool dep = lane < NB + narm;
      if (lane < D) {
        stage[lane] = dep ? pk[0] + (zk[1] * rz - zk[2] * ry) : 0.0;
        stage[D + lane] = dep ? pk[1] + (zk[2] * rx - zk[0] * rz) : 0.0;
        stage[2 * D + lane] = dep ? pk[2] + (zk[0] * ry - zk[1] * rx) : 0.0;
      }
      __syncwarp();
      for (int m = lane; m < T; m += 32) {
        int pr = 0;
        while ((pr + 1) * (pr + 2) / 2 <= m) pr++;
        const int q = m - pr * (pr + 1) / 2;
        const double h = fma(stage[2 * D + pr], stage[2 * D + q], fma(stage[D + pr], stage[D + q], stage[pr] * stage[q]));
        Hd[(N - 1) * BD + m] = fma(st.goal_w, h, Hd[(N - 1) * BD + m]);   // packed lower: entry (pr, q) of the x-x block is m
      }
      if (lane < D) {
        const double je = fma(stage[2 * D + lane], e2, fma(stage[D + lane], e1, stage[lane] * e0));
        g[last + lane] = fma(st.goal_w, je, g[last + lane]);
      }
      __syncwarp();
    }
    return lane == 0 ? 0.5 * st.goal_w * fma(e2, e2, fma(e1, e1, e0 * e0)) : 0.0;
  }
  template <bool CAND, bool GRAD>
  __device__ __forceinline__ double goal_pass() {
    const int last = (N - 1) * b;
    return goal_eval<0, GRAD>([&](int k) { return sv<CAND>(last + k); });
  }

  // ---- optional self-collision factor on every support state (gpmp2b_setting.self_collision_data):
  //      SelfCollision::evaluateError (obstacle/SelfCollision.h:66-128): e_p = hinge(eps_p - |c_A - c_B|) per sphere
  //      pair, Diagonal sigmas.  One lane evaluates support state i (sf(k) = its coordinate k): one chain walk keeps the
  //      joint lines in registers and the sphere centres in local memory, then per active pair
  //      row_k = -(n . dc_A/dq_k - n . dc_B/dq_k) / sigma  with  n . dc/dq_k = z_k . (c x n) + m_k . n  for the joints the
  //      sphere's link depends on, n = (c_A - c_B) / dist.  GRAD: the lane adds sum row row^T to the position block of
  //      ITS state's Hd and sum row e to g (no conflicts).  Returns the error share. ----
  template <int KIND, bool GRAD, class SF>
  __device__ GPMP2B_EXTRA_FN double self_eval(int i, const SF& sf) {
    constexpr int NB = (KIND == 1) ? 3 : 0;
    double eacc = 0.0;
    double zax[D][3], mom[D][3], ctr[3 * KP_MAX_SPHERES];
    double X[3], Y[3], Z[3], o[3];
    int s = 0;
    auto spheres = [&](int se) {
#pragma unroll 1
      for (; s < se; s++) {
        const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
#pragma unroll
        for (int k = 0; k < 3; k++) ctr[3 * s + k] = fma(Z[k], cz, fma(Y[k], cy, fma(X[k], cx, o[k])));
      }
    };
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < D; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) { zax[k][c] = 0.0; mom[k][c] = 0.0; }
    }
    chain_base<KIND>(sf, X, Y, Z, o, [&]() {
      if (GRAD) {
#pragma unroll
        for (int k = 0; k < 3; k++) { mom[0][k] = X[k]; mom[1 % D][k] = Y[k]; }
        zax[2 % D][2] = 1.0;
        mom[2 % D][0] = o[1]; mom[2 % D][1] = -o[0]; mom[2 % D][2] = 0.0;    // o x z
      }
      spheres(rb.sph_begin[1]);   // spheres on the vehicle (link 0)
    });
#pragma unroll 1
    for (int j = 0; j < D - NB; j++) {
      if (GRAD) {
        const double m0 = o[1] * Z[2] - o[2] * Z[1], m1 = o[2] * Z[0] - o[0] * Z[2], m2 = o[0] * Z[1] - o[1] * Z[0];
#pragma unroll
        for (int k = NB; k < D; k++)
          if (k == NB + j) {
            zax[k][0] = Z[0]; zax[k][1] = Z[1]; zax[k][2] = Z[2];
            mom[k][0] = m0; mom[k][1] = m1; mom[k][2] = m2;
          }
      }
      chain_step(j, sf(NB + j), X, Y, Z, o);
      spheres(rb.sph_begin[(KIND == 1 ? j + 1 : j) + 1]);
    }
    double M[T], cv[D];
    if (GRAD) {
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
    }
#pragma unroll 1
    for (int p = 0; p < st.n_self; p++) {
      const int sa = st.self_a[p], sb = st.self_b[p];
      const double ax = ctr[3 * sa], ay = ctr[3 * sa + 1], az = ctr[3 * sa + 2];
      const double bx = ctr[3 * sb], by = ctr[3 * sb + 1], bz = ctr[3 * sb + 2];
      const double dx = ax - bx, dy = ay - by, dz = az - bz;
      const double dist = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
      if (dist > st.self_eps[p]) continue;          // SelfCollision.h:111-117
      const double ew = (st.self_eps[p] - dist) * st.self_isig[p];
      eacc = fma(0.5 * ew, ew, eacc);
      if (GRAD) {
        const double nx = dx / dist, ny = dy / dist, nz = dz / dist;
        const double tax = ay * nz - az * ny, tay = az * nx - ax * nz, taz = ax * ny - ay * nx;   // c_A x n
        const double tbx = by * nz - bz * ny, tby = bz * nx - bx * nz, tbz = bx * ny - by * nx;   // c_B x n
        const int nja = rb.sph_link[sa] + (KIND == 1 ? 3 : 1), njb = rb.sph_link[sb] + (KIND == 1 ? 3 : 1);
        double row[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
          const double mn = fma(mom[k][2], nz, fma(mom[k][1], ny, mom[k][0] * nx));
          const double ra = k < nja ? fma(zax[k][2], taz, fma(zax[k][1], tay, fma(zax[k][0], tax, mn))) : 0.0;
          const double rb_ = k < njb ? fma(zax[k][2], tbz, fma(zax[k][1], tby, fma(zax[k][0], tbx, mn))) : 0.0;
          row[k] = (rb_ - ra) * st.self_isig[p];
        }
#pragma unroll
        for (int a = 0; a < D; a++) {
          cv[a] = fma(row[a], ew, cv[a]);
#pragma unroll
          for (int c = 0; c <= a; c++) M[a * (a + 1) / 2 + c] = fma(row[a], row[c], M[a * (a + 1) / 2 + c]);
        }
      }
    }
    if (GRAD) {
#pragma unroll
      for (int m = 0; m < T; m++) Hd[i * BD + m] += M[m];     // packed lower: the x-x block's entry m is entry m of the block
#pragma unroll
      for (int d = 0; d < D; d++) g[i * b + d] += cv[d];
    }
    return eacc;
  }
  // Rot3::Logmap of a row-major rotation matrix and SO3::LogmapDerivative, as in GTSAM 4.0.x (the same
  // branches as the oracle's restatement, which the reference's workspace-prior test vectors pin)
  static __device__ __forceinline__ void rot3_logmap(const double (&E)[9], double (&w)[3]) {
    const double tr = E[0] + E[4] + E[8];
    if (tr + 1.0 < 1e-10) {
      if (fabs(E[8] + 1.0) > 1e-5) {
        const double f = 3.14159265358979323846 / sqrt(2.0 + 2.0 * E[8]);
        w[0] = f * E[2]; w[1] = f * E[5]; w[2] = f * (1.0 + E[8]);
      } else if (fabs(E[4] + 1.0) > 1e-5) {
        const double f = 3.14159265358979323846 / sqrt(2.0 + 2.0 * E[4]);
        w[0] = f * E[1]; w[1] = f * (1.0 + E[4]); w[2] = f * E[7];
      } else {
        const double f = 3.14159265358979323846 / sqrt(2.0 + 2.0 * E[0]);
        w[0] = f * (1.0 + E[0]); w[1] = f * E[3]; w[2] = f * E[6];
      }
    } else {
      const double tr_3 = tr - 3.0;
      double magnitude;
      if (tr_3 < -1e-7) {
        const double theta = acos((tr - 1.0) / 2.0);
        magnitude = theta / (2.0 * sin(theta));
      } else {
        magnitude = 0.5 - tr_3 * tr_3 / 12.0;
      }
      w[0] = magnitude * (E[7] - E[5]); w[1] = magnitude * (E[2] - E[6]); w[2] = magnitude * (E[3] - E[1]);
    }
  }
  // L = I + W/2 + c W^2, W = skew(w); W^2 = w w^T - |w|^2 I
  static __device__ __forceinline__ void rot3_logmap_derivative(const double (&w)[3], double (&L)[9]) {
    const double t2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0]));
    double c2 = 0.0;
    if (t2 > 2.220446049250313e-16) {
      const double t = sqrt(t2);
      double sn, cs;
      sincos(t, &sn, &cs);
      c2 = 1.0 / t2 - (1.0 + cs) / (2.0 * t * sn);
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) L[r * 3 + c] = c2 * (w[r] * w[c] - (r == c ? t2 : 0.0)) + (r == c ? 1.0 : 0.0);
    L[1] -= 0.5 * w[2]; L[2] += 0.5 * w[1];
    L[3] += 0.5 * w[2]; L[5] -= 0.5 * w[0];
    L[6] -= 0.5 * w[1]; L[7] += 0.5 * w[0];
  }
  static __device__ __forceinline__ void mat33(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) C[r * 3 + c] = fma(A[r * 3 + 2], B[6 + c], fma(A[r * 3 + 1], B[3 + c], A[r * 3] * B[c]));
  }
  static __device__ __forceinline__ void skew33(const double (&v)[3], double (&W)[9]) {
    W[0] = 0.0; W[1] = -v[2]; W[2] = v[1]; W[3] = v[2]; W[4] = 0.0; W[5] = -v[0]; W[6] = -v[1]; W[7] = v[0]; W[8] = 0.0;
  }

  // ---- optional 6-D workspace pose goal on x_T (gpmp2b_setting.goal_enabled = 2):
  //      GaussianPriorWorkspacePose::evaluateError (kinematics/GaussianPriorWorkspacePose.h:53-70):
  //      e = Logmap(goal^-1 T_link) = [omega; u], H = Pose3::LogmapDerivative(e) * (body pose Jacobian), whose column k is
  //      [R^T z_k ; R^T (z_k x p + m_k)].  Pose3::Logmap / LogmapDerivative (with computeQforExpmapDerivative) as in
  //      GTSAM 4.0.x.  One lane evaluates it for support state i = N - 1; flush as self_eval. ----
  template <int KIND, bool GRAD, class SF>
  __device__ GPMP2B_EXTRA_FN double pose_eval(int i, const SF& sf) {
    constexpr int NB = (KIND == 1) ? 3 : 0;
    double zax[D][3], mom[D][3];
    double X[3], Y[3], Z[3], o[3];
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < D; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) { zax[k][c] = 0.0; mom[k][c] = 0.0; }
    }
    double vX[3] = {1.0, 0.0, 0.0}, vY[3] = {0.0, 1.0, 0.0}, vo[3] = {0.0, 0.0, 0.0};
    chain_base<KIND>(sf, X, Y, Z, o, [&]() {
      if (GRAD) {
#pragma unroll
        for (int k = 0; k < 3; k++) { mom[0][k] = X[k]; mom[1 % D][k] = Y[k]; }
        zax[2 % D][2] = 1.0;
        mom[2 % D][0] = o[1]; mom[2 % D][1] = -o[0]; mom[2 % D][2] = 0.0;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) { vX[k] = X[k]; vY[k] = Y[k]; vo[k] = o[k]; }
    });
    const int narm = (KIND == 1) ? st.goal_link : st.goal_link + 1;
    if (KIND == 1 && narm == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) { X[k] = vX[k]; Y[k] = vY[k]; Z[k] = (k == 2) ? 1.0 : 0.0; o[k] = vo[k]; }
    }
#pragma unroll 1
    for (int j = 0; j < narm; j++) {
      if (GRAD) {
        const double m0 = o[1] * Z[2] - o[2] * Z[1], m1 = o[2] * Z[0] - o[0] * Z[2], m2 = o[0] * Z[1] - o[1] * Z[0];
#pragma unroll
        for (int k = NB; k < D; k++)
          if (k == NB + j) {
            zax[k][0] = Z[0]; zax[k][1] = Z[1]; zax[k][2] = Z[2];
            mom[k][0] = m0; mom[k][1] = m1; mom[k][2] = m2;
          }
      }
      chain_step(j, sf(NB + j), X, Y, Z, o);
    }
    // E = Rd^T R, te = Rd^T (p - td)
    double E[9], te[3], xi[6];
    const double dp[3] = {o[0] - goal_pos_p[0], o[1] - goal_pos_p[1], o[2] - goal_pos_p[2]};
#pragma unroll
    for (int r = 0; r < 3; r++) {
      E[r * 3 + 0] = fma(goal_R_p[6 + r], X[2], fma(goal_R_p[3 + r], X[1], goal_R_p[r] * X[0]));
      E[r * 3 + 1] = fma(goal_R_p[6 + r], Y[2], fma(goal_R_p[3 + r], Y[1], goal_R_p[r] * Y[0]));
      E[r * 3 + 2] = fma(goal_R_p[6 + r], Z[2], fma(goal_R_p[3 + r], Z[1], goal_R_p[r] * Z[0]));
      te[r] = fma(goal_R_p[6 + r], dp[2], fma(goal_R_p[3 + r], dp[1], goal_R_p[r] * dp[0]));
    }
    double w[3];
    rot3_logmap(E, w);
    const double th2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0])), th = sqrt(th2);
    double u[3];
    if (th < 1e-10) {
      u[0] = te[0]; u[1] = te[1]; u[2] = te[2];
    } else {
      const double k0 = w[0] / th, k1 = w[1] / th, k2 = w[2] / th;
      const double a0 = k1 * te[2] - k2 * te[1], a1 = k2 * te[0] - k0 * te[2], a2 = k0 * te[1] - k1 * te[0];
      const double b0 = k1 * a2 - k2 * a1, b1 = k2 * a0 - k0 * a2, b2 = k0 * a1 - k1 * a0;
      const double cc = 1.0 - th / (2.0 * tan(0.5 * th));
      u[0] = te[0] - 0.5 * th * a0 + cc * b0;
      u[1] = te[1] - 0.5 * th * a1 + cc * b1;
      u[2] = te[2] - 0.5 * th * a2 + cc * b2;
    }
    xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2]; xi[3] = u[0]; xi[4] = u[1]; xi[5] = u[2];
    const double eacc = 0.5 * st.goal_w * (th2 + fma(u[2], u[2], fma(u[1], u[1], u[0] * u[0])));
    if (GRAD) {
      double Jw[9], Wm[9], Vm[9], WV[9], VW[9], WVW[9], WW[9], T1[9], T2[9], Q[9], Q2[9];
      rot3_logmap_derivative(w, Jw);
      skew33(w, Wm); skew33(u, Vm);
      mat33(Wm, Vm, WV); mat33(Vm, Wm, VW); mat33(WV, Wm, WVW); mat33(Wm, Wm, WW);
      double c1, c2, c3;
      if (th > 1e-5) {
        double sn, cs;
        sincos(th, &sn, &cs);
        const double p3 = th2 * th, p4 = th2 * th2, p5 = p4 * th;
        c1 = (th - sn) / p3;
        c2 = (1.0 - 0.5 * th2 - cs) / p4;
        c3 = -0.5 * (c2 - 3.0 * (th - sn - p3 / 6.0) / p5);
      } else {
        c1 = 1.0 / 6.0; c2 = 1.0 / 24.0; c3 = -0.5 * (1.0 / 24.0 + 3.0 / 120.0);
      }
      mat33(WW, Vm, T1);       // WWV
      mat33(VW, Wm, T2);       // VWW
#pragma unroll
      for (int e = 0; e < 9; e++) Q[e] = -0.5 * Vm[e] + c1 * (WV[e] + VW[e] - WVW[e]) + c2 * (T1[e] + T2[e] - 3.0 * WVW[e]);
      mat33(WVW, Wm, T1);      // WVWW
      mat33(Wm, WVW, T2);      // WWVW
#pragma unroll
      for (int e = 0; e < 9; e++) Q[e] += c3 * (T1[e] + T2[e]);
      mat33(Jw, Q, T1); mat33(T1, Jw, Q2);
#pragma unroll
      for (int e = 0; e < 9; e++) Q2[e] = -Q2[e];
      const int nj = NB + narm;
      const double isig = sqrt(st.goal_w);
      double M[T], cv[D];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
      double ab[D][3], lb[D][3];   // body twist per joint: [R^T z_k ; R^T (z_k x p + m_k)]
#pragma unroll
      for (int k = 0; k < D; k++) {
        const bool dep = k < nj;
        const double l0 = zax[k][1] * o[2] - zax[k][2] * o[1] + mom[k][0], l1 = zax[k][2] * o[0] - zax[k][0] * o[2] + mom[k][1],
                     l2 = zax[k][0] * o[1] - zax[k][1] * o[0] + mom[k][2];
        ab[k][0] = dep ? fma(X[2], zax[k][2], fma(X[1], zax[k][1], X[0] * zax[k][0])) : 0.0;
        ab[k][1] = dep ? fma(Y[2], zax[k][2], fma(Y[1], zax[k][1], Y[0] * zax[k][0])) : 0.0;
        ab[k][2] = dep ? fma(Z[2], zax[k][2], fma(Z[1], zax[k][1], Z[0] * zax[k][0])) : 0.0;
        lb[k][0] = dep ? fma(X[2], l2, fma(X[1], l1, X[0] * l0)) : 0.0;
        lb[k][1] = dep ? fma(Y[2], l2, fma(Y[1], l1, Y[0] * l0)) : 0.0;
        lb[k][2] = dep ? fma(Z[2], l2, fma(Z[1], l1, Z[0] * l0)) : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 6; r++) {
        double row[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
          double v;
          if (r < 3) v = fma(Jw[r * 3 + 2], ab[k][2], fma(Jw[r * 3 + 1], ab[k][1], Jw[r * 3] * ab[k][0]));
          else {
            const int q = r - 3;
            v = fma(Q2[q * 3 + 2], ab[k][2], fma(Q2[q * 3 + 1], ab[k][1], Q2[q * 3] * ab[k][0]));
            v = fma(Jw[q * 3 + 2], lb[k][2], fma(Jw[q * 3 + 1], lb[k][1], fma(Jw[q * 3], lb[k][0], v)));
          }
          row[k] = v * isig;
        }
        const double ew = xi[r] * isig;
#pragma unroll
        for (int a = 0; a < D; a++) {
          cv[a] = fma(row[a], ew, cv[a]);
#pragma unroll
          for (int c = 0; c <= a; c++) M[a * (a + 1) / 2 + c] = fma(row[a], row[c], M[a * (a + 1) / 2 + c]);
        }
      }
#pragma unroll
      for (int m = 0; m < T; m++) Hd[i * BD + m] += M[m];
#pragma unroll
      for (int d = 0; d < D; d++) g[i * b + d] += cv[d];
    }
    return eacc;
  }

  // ---- optional workspace orientation prior on support states orient_first..orient_last (gpmp2b_setting.orient_*):
  //      GaussianPriorWorkspaceOrientation::evaluateError (kinematics/GaussianPriorWorkspaceOrientation.h:53-72):
  //      e = Logmap(des^T R_link), H = LogmapDerivative(e) * (body angular velocity Jacobian) -- column k of the latter is
  //      R^T z_k for the joints the link depends on (the rows [I 0] * J_jpx_jp picks).  Rot3 Logmap / LogmapDerivative as
  //      in GTSAM.  One lane evaluates support state i; flush as self_eval. ----
  template <int KIND, bool GRAD, class SF>
  __device__ GPMP2B_EXTRA_FN double orient_eval(int i, const SF& sf) {
    constexpr int NB = (KIND == 1) ? 3 : 0;
    double zax[D][3];
    double X[3], Y[3], Z[3], o[3];
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < D; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) zax[k][c] = 0.0;
    }
    double vX[3] = {1.0, 0.0, 0.0}, vY[3] = {0.0, 1.0, 0.0};
    chain_base<KIND>(sf, X, Y, Z, o, [&]() {
      if (GRAD) zax[2 % D][2] = 1.0;
#pragma unroll
      for (int k = 0; k < 3; k++) { vX[k] = X[k]; vY[k] = Y[k]; }
    });
    const int narm = (KIND == 1) ? st.orient_link : st.orient_link + 1;
    if (KIND == 1 && narm == 0) {   // link 0 = the vehicle frame itself
#pragma unroll
      for (int k = 0; k < 3; k++) { X[k] = vX[k]; Y[k] = vY[k]; Z[k] = (k == 2) ? 1.0 : 0.0; }
    }
#pragma unroll 1
    for (int j = 0; j < narm; j++) {
      if (GRAD) {
#pragma unroll
        for (int k = NB; k < D; k++)
          if (k == NB + j) { zax[k][0] = Z[0]; zax[k][1] = Z[1]; zax[k][2] = Z[2]; }
      }
      chain_step(j, sf(NB + j), X, Y, Z, o);
    }
    // E = des^T R, R = [X Y Z]
    double E[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      E[r * 3 + 0] = fma(orient_R_p[6 + r], X[2], fma(orient_R_p[3 + r], X[1], orient_R_p[r] * X[0]));
      E[r * 3 + 1] = fma(orient_R_p[6 + r], Y[2], fma(orient_R_p[3 + r], Y[1], orient_R_p[r] * Y[0]));
      E[r * 3 + 2] = fma(orient_R_p[6 + r], Z[2], fma(orient_R_p[3 + r], Z[1], orient_R_p[r] * Z[0]));
    }
    double w[3];
    rot3_logmap(E, w);
    const double t2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0]));
    const double eacc = 0.5 * st.orient_w * t2;
    if (GRAD) {
      double L[9];
      rot3_logmap_derivative(w, L);
      const int nj = NB + narm;
      const double isig = sqrt(st.orient_w);
      double M[T], cv[D];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
      double ab[D][3];   // body angular velocity per joint: R^T z_k
#pragma unroll
      for (int k = 0; k < D; k++) {
        const bool dep = k < nj;
        ab[k][0] = dep ? fma(X[2], zax[k][2], fma(X[1], zax[k][1], X[0] * zax[k][0])) : 0.0;
        ab[k][1] = dep ? fma(Y[2], zax[k][2], fma(Y[1], zax[k][1], Y[0] * zax[k][0])) : 0.0;
        ab[k][2] = dep ? fma(Z[2], zax[k][2], fma(Z[1], zax[k][1], Z[0] * zax[k][0])) : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double row[D];
#pragma unroll
        for (int k = 0; k < D; k++) row[k] = fma(L[r * 3 + 2], ab[k][2], fma(L[r * 3 + 1], ab[k][1], L[r * 3] * ab[k][0])) * isig;
        const double ew = w[r] * isig;
#pragma unroll
        for (int a = 0; a < D; a++) {
          cv[a] = fma(row[a], ew, cv[a]);
#pragma unroll
          for (int c = 0; c <= a; c++) M[a * (a + 1) / 2 + c] = fma(row[a], row[c], M[a * (a + 1) / 2 + c]);
        }
      }
#pragma unroll
      for (int m = 0; m < T; m++) Hd[i * BD + m] += M[m];
#pragma unroll
      for (int d = 0; d < D; d++) g[i * b + d] += cv[d];
    }
    return eacc;
  }
  template <bool CAND, bool GRAD>
  __device__ __forceinline__ double orient_pass() {
    double eacc = 0.0;
#pragma unroll 1
    for (int i = st.orient_first + lane; i <= st.orient_last; i += 32)
      eacc += orient_eval<0, GRAD>(i, [&](int k) { return sv<CAND>(i * b + k); });
    return eacc;
  }
  template <bool CAND, bool GRAD>
  __device__ __forceinline__ double self_pass() {
    double eacc = 0.0;
#pragma unroll 1
    for (int i = lane; i < N; i += 32) eacc += self_eval<0, GRAD>(i, [&](int k) { return sv<CAND>(i * b + k); });
    return eacc;
  }

  // The H storage (Ho | Hd, contiguous) is dead whenever an error is evaluated -- before the first linearization,
  // after the solve (LM restores H from its backup on a rejected step, an accepted step re-linearizes) -- so the
  // error pass borrows it as the landing zone of its asynchronous SDF gathers (device_model.cuh: config_error).
  // 3 KB per sphere of a chunk; nullptr (register path) when it does not hold at least two spheres.
  __device__ __forceinline__ double* err_scratch(int& chunk) const {
    if (escr_chunk >= 2) { chunk = escr_chunk; return escr; }
    chunk = no_err_scratch ? 0 : min(8, (N * BD + (N - 1) * BB - err_scratch_off) / 384);
    return chunk >= 2 ? Ho + err_scratch_off : nullptr;
  }

  // ---- NonlinearFactorGraph::error at xs (CAND=false) or xs+dl (CAND=true) ----
  // ---- PriorFactor(x_k, fix_conf) + PriorFactor(v_k, fix_vel) on support state k = st.fix_index: the pinned state of a
  //      replanning re-solve (ISAM2TrajOptimizer::fixConfigAndVel, ISAM2TrajOptimizer-inl.h:160-168).  Returns this lane's
  //      share of the error; GRAD: adds the gradient (the constant Hessian is in the host's template). ----
  template <bool CAND, bool GRAD>
  __device__ __forceinline__ double fix_pass() {
    const int i = st.fix_index;
    double e = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double ex = sv<CAND>(i * b + d) - fix_conf[d], ev = sv<CAND>(i * b + D + d) - fix_vel[d];
      if (GRAD) { g[i * b + d] += st.conf_prior_w * ex; g[i * b + D + d] += st.vel_prior_w * ev; }
      e += 0.5 * (st.conf_prior_w * ex * ex + st.vel_prior_w * ev * ev);
    }
    return e;
  }

  // MASK (error kernel of the phase pipeline): also record, per configuration, which spheres are within reach of their
  // hinge (mask_out[cidx], one bit per sphere) for the linearization that follows an accepted step at the same states
  template <bool CAND, bool MASK = false>
  __device__ double eval_error() {
    double eacc = state_pass<CAND, false>();
    if constexpr (EXTRA) {
      if ((GPMP2B_EXTRA_MASK & 1) && st.goal_enabled == 1) eacc += goal_pass<CAND, false>();
      if ((GPMP2B_EXTRA_MASK & 8) && st.goal_enabled == 2 && lane == 0)
        eacc += pose_eval<0, false>(N - 1, [&](int k) { return sv<CAND>((N - 1) * b + k); });
      if ((GPMP2B_EXTRA_MASK & 2) && st.n_self) eacc += self_pass<CAND, false>();
      if ((GPMP2B_EXTRA_MASK & 4) && st.orient_enabled) eacc += orient_pass<CAND, false>();
      if (st.fix_enabled) eacc += fix_pass<CAND, false>();
    }
    int chunk;
    double* scratch = err_scratch(chunk);
    __syncwarp();
    double e2 = 0.0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      if (cidx < C) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double es = 0.0;
        if constexpr (MASK) {
          unsigned long long am = 0ull;
          config_error<D, NDIM, 0, false, true>(rb, sdf, config_state<CAND>(i, j), st.epsilon, st.inv_cost_sigma, e2, es, nullptr, nullptr,
                                                nullptr, 0, &am);
          mask_out[cidx] = am;
        } else {
          config_error<D, NDIM, 0, false>(rb, sdf, config_state<CAND>(i, j), st.epsilon, st.inv_cost_sigma, e2, es, nullptr, nullptr,
                                          scratch, chunk);
        }
      }
    }
    return warp_sum(eacc + 0.5 * e2);
  }

  // ---- CollisionCost (BatchTrajOptimizer-inl.h:87-100): eps = 0, unwhitened sum over support states ----
  __device__ double collision_cost() {
    double es = 0.0;
    int chunk;
    double* scratch = err_scratch(chunk);
    for (int i = lane; i < N; i += 32) {
      double e2 = 0.0;
      config_error<D, NDIM, 0, false>(rb, sdf, config_state<false>(i, 0), 0.0, 1.0, e2, es, nullptr, nullptr, scratch, chunk);
    }
    return warp_sum(es);
  }

  // ---- NonlinearFactorGraph::linearize folded straight into the block-tridiagonal normal equations ----
  // everything of the linearization except the obstacle factors: constant template, per-state pass, optional factors
  __device__ __forceinline__ void linearize_head() {
    // constant part of H: GP-prior blocks + end-state priors (host-precomputed template, same layout as smem)
    {
      copy_in(reinterpret_cast<const double2*>(hconst), reinterpret_cast<double2*>(Ho), (N * BD + (N - 1) * BB + 1) / 2);
    }
    __syncwarp();
    state_pass<false, true>();
    __syncwarp();
    if constexpr (EXTRA) {
      if ((GPMP2B_EXTRA_MASK & 1) && st.goal_enabled == 1) goal_pass<false, true>();
      if ((GPMP2B_EXTRA_MASK & 8) && st.goal_enabled == 2) {
        if (lane == 0) pose_eval<0, true>(N - 1, [&](int k) { return xs[(N - 1) * b + k]; });
        __syncwarp();
      }
      if ((GPMP2B_EXTRA_MASK & 2) && st.n_self) { self_pass<false, true>(); __syncwarp(); }
      if ((GPMP2B_EXTRA_MASK & 4) && st.orient_enabled) { orient_pass<false, true>(); __syncwarp(); }
      if (st.fix_enabled) { fix_pass<false, true>(); __syncwarp(); }
    }
  }

  // ---- phase-kernel pipeline, linearize kernel: the configuration-parallel half of the linearization.  Every lane
  //      evaluates configurations c = lane, lane + 32, ... and hands its whitened (M, cv) to the global M-list,
  //      configuration-major (ml[c * RS + m], RS = T + D padded to even): the entry-parallel lanes of the solve kernel
  //      then read entry m of one configuration with ONE coalesced request (the transposed layout cost the solve
  //      kernel 28 wavefronts per load -- a quarter of its LSU pipe).  The rows go through shared memory (stg: 32 rows)
  //      so that the stores coalesce too. ----
  __device__ void linearize_configs_to_global(double* __restrict__ ml, int RS, double* stg, int gather_chunk = 0) {
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      double M[T], cv[D];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
      if (cidx < C) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double e2 = 0.0, es = 0.0;
        // 3-D fields: asynchronous gathers through the staging buffer (dead until the rows are staged)
#if GPMP2B_LIN_ASYNC   // (compile-time: both paths in one kernel cost the default one 20 registers)
        if (NDIM == 3 && gather_chunk >= 2)
          config_eval_async<D>(rb, sdf, config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, M, cv, e2, es, stg, gather_chunk);
        else
#endif
          config_eval<D, NDIM, 0, true, false>(rb, sdf, config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, M, cv, e2, es,
                                               nullptr, nullptr);
      }
      __syncwarp();                                    // every lane is done with its gather slots
      if (cidx < C) {
        double* o = stg + lane * RS;
#pragma unroll
        for (int m = 0; m < T; m++) o[m] = M[m];
#pragma unroll
        for (int d = 0; d < D; d++) o[T + d] = cv[d];
      }
      __syncwarp();
      // the pass's rows are contiguous in the M-list: one coalesced copy
      const int n = min(32, C - c0) * RS;
      double* dst = ml + (size_t)c0 * RS;
      for (int idx = lane; idx < n; idx += 32) dst[idx] = stg[idx];
      __syncwarp();
    }
  }

  // ---- phase-kernel pipeline, solve kernel: the entry-parallel half.  Same arithmetic in the same order as
  //      linearize() (bit-identical H and g): the staging buffer is replaced by this lane's rows of the M-list. ----
  __device__ void assemble_from_global(const double* __restrict__ ml, int RS) {
    linearize_head();
#if GPMP2B_ALIGNED_ACC
    if (K == 5) { assemble_obstacles_aligned<5>(ml, RS); return; }
    if (K == 4) { assemble_obstacles_aligned<4>(ml, RS); return; }
#endif
    assemble_obstacles_generic(ml, RS);
  }

  __device__ void linearize() {
    linearize_head();

#if GPMP2B_ALIGNED_ACC
    if (K == 5) { linearize_obstacles_aligned<5>(); return; }   // the library default ...
    if (K == 4) { linearize_obstacles_aligned<4>(); return; }   // ... and the 2-D examples' setting
#endif
    // entry-parallel phase: this lane owns packed entry m = lane = (p, q) of every symmetric D x D sub-block
    const int p = tp, q = tq;
    const int dxx = p * (p + 1) / 2 + q, dvx1 = (D + p) * (D + p + 1) / 2 + q, dvx2 = (D + q) * (D + q + 1) / 2 + p,
              dvv = (D + p) * (D + p + 1) / 2 + D + q;
    const int o1 = p * b + q, o2 = q * b + p;
    const bool offdiag = p != q, hlane = lane < T, glane = lane < D;

    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      const bool valid = cidx < C;
      double M[T], cv[D];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
#ifdef GPMP2B_PHASE_TIMING
      long long tc0 = clock64();
#endif
      if (valid) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double e2 = 0.0, es = 0.0;
        config_eval<D, NDIM, 0, true, false>(rb, sdf, config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, M, cv, e2, es,
                                             nullptr, nullptr);
      }
#ifdef GPMP2B_PHASE_TIMING
      __syncwarp();
      pt_cfg += clock64() - tc0;
      tc0 = clock64();
#endif
      // hand the per-configuration (M, cv) to the entry-parallel lanes, 8 configurations per round
      int ri = c0 / (K + 1), rj = c0 - ri * (K + 1);   // (i, j) of the first configuration of the round
#pragma unroll 1
      for (int round = 0; round < 4; round++) {
        const int ci0 = c0 + round * 8;
        if (ci0 >= C) break;
        if ((lane >> 3) == round && valid) {
          double* sp = stage + (lane & 7) * STG;
#pragma unroll
          for (int m = 0; m < T; m++) sp[m] = M[m];
#pragma unroll
          for (int d = 0; d < D; d++) sp[T + d] = cv[d];
        }
        __syncwarp();
        const int nt = min(8, C - ci0);
        int t = 0;
#pragma unroll 1
        while (t < nt) {
          // one segment = the configurations of interval ri present in this round
          double a0xx = 0, a0xv = 0, a0vv = 0, a1xx = 0, a1xv = 0, a1vv = 0, oxx = 0, oxv = 0, ovx = 0, ovv = 0;
          double g0x = 0, g0v = 0, g1x = 0, g1v = 0;
          const double* sp = stage + t * STG + lane;
          if (rj == 0) {   // unary factor on x_i
            if (hlane) a0xx = sp[0];
            if (glane) g0x = sp[T];
            sp += STG; t++; rj = 1;
          }
          const int jn = min(K + 1 - rj, nt - t);   // GP obstacle factors of this segment
#pragma unroll 1
          for (int u = 0; u < jn; u++) {
            // H += (w w^T) (x) M on (x_i, v_i, x_{i+1}, v_{i+1})
            const double val = hlane ? sp[0] : 0.0;
            const double gval = glane ? sp[T] : 0.0;
            const double* ww = st.gpww[rj - 1 + u];
            const double* w = st.gpw[rj - 1 + u];
            a0xx = fma(ww[0], val, a0xx); a0xv = fma(ww[1], val, a0xv); a0vv = fma(ww[2], val, a0vv);
            oxx = fma(ww[3], val, oxx);   oxv = fma(ww[4], val, oxv);
            ovx = fma(ww[5], val, ovx);   ovv = fma(ww[6], val, ovv);
            a1xx = fma(ww[7], val, a1xx); a1xv = fma(ww[8], val, a1xv); a1vv = fma(ww[9], val, a1vv);
            g0x = fma(w[0], gval, g0x); g0v = fma(w[1], gval, g0v);
            g1x = fma(w[2], gval, g1x); g1v = fma(w[3], gval, g1v);
            sp += STG;
          }
          t += jn; rj += jn;
          // write the segment out (every lane owns its entries: no conflicts).  Loads first, then stores: the
          // offsets are run-time values, so interleaved read-modify-writes would be serialised by the compiler.
          if (hlane) {
            double* Hdi = Hd + ri * BD;
            const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
            if (jn > 0) {
              double* Hoi = Ho + ri * BB;
              double* Hdn = Hdi + BD;
              const double q0 = Hoi[o1], q1 = Hoi[o1 + D], q2 = Hoi[o1 + D * b], q3 = Hoi[o1 + D * b + D];
              const double q4 = Hoi[o2], q5 = Hoi[o2 + D], q6 = Hoi[o2 + D * b], q7 = Hoi[o2 + D * b + D];
              const double n0 = Hdn[dxx], n1 = Hdn[dvx1], n2 = Hdn[dvx2], n3 = Hdn[dvv];
              Hoi[o1] = q0 + oxx;               Hoi[o1 + D] = q1 + oxv;
              Hoi[o1 + D * b] = q2 + ovx;       Hoi[o1 + D * b + D] = q3 + ovv;
              if (offdiag) {
                Hoi[o2] = q4 + oxx;             Hoi[o2 + D] = q5 + oxv;
                Hoi[o2 + D * b] = q6 + ovx;     Hoi[o2 + D * b + D] = q7 + ovv;
              }
              Hdn[dxx] = n0 + a1xx;
              Hdn[dvx1] = n1 + a1xv;
              if (offdiag) Hdn[dvx2] = n2 + a1xv;
              Hdn[dvv] = n3 + a1vv;
            }
            Hdi[dxx] = h0 + a0xx;
            Hdi[dvx1] = h1 + a0xv;
            if (offdiag) Hdi[dvx2] = h2 + a0xv;
            Hdi[dvv] = h3 + a0vv;
          }
          if (glane) {
            double* gi = g + ri * b + lane;
            const double t0 = gi[0], t1 = gi[D];
            if (jn > 0) {
              const double t2 = gi[b], t3 = gi[b + D];
              gi[b] = t2 + g1x;
              gi[b + D] = t3 + g1v;
            }
            gi[0] = t0 + g0x;
            gi[D] = t1 + g0v;
          }
          if (rj > K) { rj = 0; ri++; }
        }
        __syncwarp();
      }
#ifdef GPMP2B_PHASE_TIMING
      pt_acc += clock64() - tc0;
#endif
    }
    __syncwarp();
  }

  // ---- obstacle factors, interval-aligned variant for a compile-time obs_check_inter KS (the library default 5):
  //      a pass evaluates 32 / (KS + 1) whole intervals, so an interval's unary + KS interpolated factors are staged
  //      and reduced together with STATIC j (the GP weights become constant-bank operands of the DFMAs instead of 14
  //      indexed constant loads per factor), each block of H is read-modified-written once per interval, and the
  //      contribution to the next diagonal block is carried in registers into the next interval. ----
  template <int KS>
  __device__ void linearize_obstacles_aligned() {
    constexpr int CI = KS + 1, IPP = 32 / CI;
    static_assert(CI * STG <= 8 * STG, "staging buffer holds 8 configurations");
    const int p = tp, q = tq;
    const int dxx = p * (p + 1) / 2 + q, dvx1 = (D + p) * (D + p + 1) / 2 + q, dvx2 = (D + q) * (D + q + 1) / 2 + p,
              dvv = (D + p) * (D + p + 1) / 2 + D + q;
    const int o1 = p * b + q, o2 = q * b + p;
    const bool offdiag = p != q, hlane = lane < T, glane = lane < D;
    const int li = lane / CI, lj = lane - li * CI;
#pragma unroll 1
    for (int i0 = 0; i0 < N - 1; i0 += IPP) {
      int ci = i0 + li, cj = lj;
      bool valid = li < IPP && ci < N - 1;
      const bool last_state = i0 == 0 && lane == IPP * CI;     // the unary factor of x_{N-1} rides along in pass 0
      if (last_state) { ci = N - 1; cj = 0; valid = true; }
      double M[T], cv[D];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
      if (valid) {
        double e2 = 0.0, es = 0.0;
        config_eval<D, NDIM, 0, true, false>(rb, sdf, config_state<false>(ci, cj), st.epsilon, st.inv_cost_sigma, M, cv, e2, es,
                                             nullptr, nullptr);
      }
      double cxx = 0.0, cxv = 0.0, cvv = 0.0, cgx = 0.0, cgv = 0.0;   // carried into the next diagonal block
      const int ns = min(IPP, N - 1 - i0);
#pragma unroll 1
      for (int sl = 0; sl <= ns; sl++) {
        // slot ns = the lone unary factor of the last state (pass 0 only)
        const bool tail = sl == ns;
        if (tail && i0 != 0) break;
        if (tail ? last_state : (li == sl && li < IPP)) {
          double* sp = stage + (tail ? 0 : lj) * STG;
#pragma unroll
          for (int m = 0; m < T; m++) sp[m] = M[m];
#pragma unroll
          for (int d = 0; d < D; d++) sp[T + d] = cv[d];
        }
        __syncwarp();
        const double* sp = stage + lane;
        if (tail) {
          // flush the carry of this pass's last interval, then the last state's unary factor
          if (hlane) {
            double* Hdi = Hd + (i0 + ns) * BD;
            const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
            Hdi[dxx] = h0 + cxx; Hdi[dvx1] = h1 + cxv; if (offdiag) Hdi[dvx2] = h2 + cxv; Hdi[dvv] = h3 + cvv;
            Hd[(N - 1) * BD + dxx] += sp[0];
          }
          if (glane) {
            double* gi = g + (i0 + ns) * b + lane;
            gi[0] += cgx; gi[D] += cgv;
            g[(N - 1) * b + lane] += sp[T];
          }
          cxx = cxv = cvv = cgx = cgv = 0.0;
        } else {
          const int i = i0 + sl;
          double a0xx = cxx + (hlane ? sp[0] : 0.0), a0xv = cxv, a0vv = cvv, a1xx = 0, a1xv = 0, a1vv = 0;
          double oxx = 0, oxv = 0, ovx = 0, ovv = 0;
          double g0x = cgx + (glane ? sp[T] : 0.0), g0v = cgv, g1x = 0, g1v = 0;
#pragma unroll
          for (int j = 1; j <= KS; j++) {
            const double val = hlane ? sp[j * STG] : 0.0;
            const double gval = glane ? sp[j * STG + T] : 0.0;
            a0xx = fma(st.gpww[j - 1][0], val, a0xx); a0xv = fma(st.gpww[j - 1][1], val, a0xv); a0vv = fma(st.gpww[j - 1][2], val, a0vv);
            oxx = fma(st.gpww[j - 1][3], val, oxx);   oxv = fma(st.gpww[j - 1][4], val, oxv);
            ovx = fma(st.gpww[j - 1][5], val, ovx);   ovv = fma(st.gpww[j - 1][6], val, ovv);
            a1xx = fma(st.gpww[j - 1][7], val, a1xx); a1xv = fma(st.gpww[j - 1][8], val, a1xv); a1vv = fma(st.gpww[j - 1][9], val, a1vv);
            g0x = fma(st.gpw[j - 1][0], gval, g0x); g0v = fma(st.gpw[j - 1][1], gval, g0v);
            g1x = fma(st.gpw[j - 1][2], gval, g1x); g1v = fma(st.gpw[j - 1][3], gval, g1v);
          }
          if (hlane) {
            double* Hdi = Hd + i * BD;
            double* Hoi = Ho + i * BB;
            const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
            const double q0 = Hoi[o1], q1 = Hoi[o1 + D], q2 = Hoi[o1 + D * b], q3 = Hoi[o1 + D * b + D];
            const double q4 = Hoi[o2], q5 = Hoi[o2 + D], q6 = Hoi[o2 + D * b], q7 = Hoi[o2 + D * b + D];
            Hoi[o1] = q0 + oxx;               Hoi[o1 + D] = q1 + oxv;
            Hoi[o1 + D * b] = q2 + ovx;       Hoi[o1 + D * b + D] = q3 + ovv;
            if (offdiag) {
              Hoi[o2] = q4 + oxx;             Hoi[o2 + D] = q5 + oxv;
              Hoi[o2 + D * b] = q6 + ovx;     Hoi[o2 + D * b + D] = q7 + ovv;
            }
            Hdi[dxx] = h0 + a0xx;
            Hdi[dvx1] = h1 + a0xv;
            if (offdiag) Hdi[dvx2] = h2 + a0xv;
            Hdi[dvv] = h3 + a0vv;
          }
          if (glane) {
            double* gi = g + i * b + lane;
            const double t0 = gi[0], t1 = gi[D];
            gi[0] = t0 + g0x;
            gi[D] = t1 + g0v;
          }
          cxx = a1xx; cxv = a1xv; cvv = a1vv; cgx = g1x; cgv = g1v;
        }
        __syncwarp();
      }
      if (i0 != 0 && hlane) {      // carry of the last interval of a later pass
        double* Hdi = Hd + (i0 + ns) * BD;
        const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
        Hdi[dxx] = h0 + cxx; Hdi[dvx1] = h1 + cxv; if (offdiag) Hdi[dvx2] = h2 + cxv; Hdi[dvv] = h3 + cvv;
      }
      if (i0 != 0 && glane) {
        double* gi = g + (i0 + ns) * b + lane;
        gi[0] += cgx; gi[D] += cgv;
      }
      __syncwarp();
    }
  }

  // ---- assemble_obstacles_*: linearize_obstacles_* with the per-configuration (M, cv) read from the global M-list
  //      (this lane's entry row rowM = ml + lane * CS, gradient row rowG = ml + (T + lane) * CS) instead of the staging
  //      buffer.  The pass / round / segment structure of the fused versions is kept only as index arithmetic, so that
  //      every sum is associated exactly as there. ----
  template <int KS>
  __device__ void assemble_obstacles_aligned(const double* __restrict__ ml, int RS) {
    constexpr int CI = KS + 1, IPP = 32 / CI;
    const int p = tp, q = tq;
    const int dxx = p * (p + 1) / 2 + q, dvx1 = (D + p) * (D + p + 1) / 2 + q, dvx2 = (D + q) * (D + q + 1) / 2 + p,
              dvv = (D + p) * (D + p + 1) / 2 + D + q;
    const int o1 = p * b + q, o2 = q * b + p;
    const bool offdiag = p != q, hlane = lane < T, glane = lane < D;
    const double* rowM = ml + (hlane ? lane : 0);          // entry `lane` of configuration c: rowM[c * RS]
    const double* rowG = ml + T + (glane ? lane : 0);
#pragma unroll 1
    for (int i0 = 0; i0 < N - 1; i0 += IPP) {
      double cxx = 0.0, cxv = 0.0, cvv = 0.0, cgx = 0.0, cgv = 0.0;   // carried into the next diagonal block
      const int ns = min(IPP, N - 1 - i0);
#pragma unroll 1
      for (int sl = 0; sl < ns; sl++) {
        const int i = i0 + sl;
        const double* vm = rowM + (size_t)(i * CI) * RS;
        const double* vg = rowG + (size_t)(i * CI) * RS;
        double val[CI], gval[CI];
#pragma unroll
        for (int j = 0; j < CI; j++) { val[j] = hlane ? __ldg(vm + j * RS) : 0.0; gval[j] = glane ? __ldg(vg + j * RS) : 0.0; }
        double a0xx = cxx + val[0], a0xv = cxv, a0vv = cvv, a1xx = 0, a1xv = 0, a1vv = 0;
        double oxx = 0, oxv = 0, ovx = 0, ovv = 0;
        double g0x = cgx + gval[0], g0v = cgv, g1x = 0, g1v = 0;
#pragma unroll
        for (int j = 1; j <= KS; j++) {
          a0xx = fma(st.gpww[j - 1][0], val[j], a0xx); a0xv = fma(st.gpww[j - 1][1], val[j], a0xv); a0vv = fma(st.gpww[j - 1][2], val[j], a0vv);
          oxx = fma(st.gpww[j - 1][3], val[j], oxx);   oxv = fma(st.gpww[j - 1][4], val[j], oxv);
          ovx = fma(st.gpww[j - 1][5], val[j], ovx);   ovv = fma(st.gpww[j - 1][6], val[j], ovv);
          a1xx = fma(st.gpww[j - 1][7], val[j], a1xx); a1xv = fma(st.gpww[j - 1][8], val[j], a1xv); a1vv = fma(st.gpww[j - 1][9], val[j], a1vv);
          g0x = fma(st.gpw[j - 1][0], gval[j], g0x); g0v = fma(st.gpw[j - 1][1], gval[j], g0v);
          g1x = fma(st.gpw[j - 1][2], gval[j], g1x); g1v = fma(st.gpw[j - 1][3], gval[j], g1v);
        }
        if (hlane) {
          double* Hdi = Hd + i * BD;
          double* Hoi = Ho + i * BB;
          const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
          const double q0 = Hoi[o1], q1 = Hoi[o1 + D], q2 = Hoi[o1 + D * b], q3 = Hoi[o1 + D * b + D];
          const double q4 = Hoi[o2], q5 = Hoi[o2 + D], q6 = Hoi[o2 + D * b], q7 = Hoi[o2 + D * b + D];
          Hoi[o1] = q0 + oxx;               Hoi[o1 + D] = q1 + oxv;
          Hoi[o1 + D * b] = q2 + ovx;       Hoi[o1 + D * b + D] = q3 + ovv;
          if (offdiag) {
            Hoi[o2] = q4 + oxx;             Hoi[o2 + D] = q5 + oxv;
            Hoi[o2 + D * b] = q6 + ovx;     Hoi[o2 + D * b + D] = q7 + ovv;
          }
          Hdi[dxx] = h0 + a0xx;
          Hdi[dvx1] = h1 + a0xv;
          if (offdiag) Hdi[dvx2] = h2 + a0xv;
          Hdi[dvv] = h3 + a0vv;
        }
        if (glane) {
          double* gi = g + i * b + lane;
          const double t0 = gi[0], t1 = gi[D];
          gi[0] = t0 + g0x;
          gi[D] = t1 + g0v;
        }
        cxx = a1xx; cxv = a1xv; cvv = a1vv; cgx = g1x; cgv = g1v;
        __syncwarp();
      }
      // the carry of the pass's last interval ...
      if (hlane) {
        double* Hdi = Hd + (i0 + ns) * BD;
        const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
        Hdi[dxx] = h0 + cxx; Hdi[dvx1] = h1 + cxv; if (offdiag) Hdi[dvx2] = h2 + cxv; Hdi[dvv] = h3 + cvv;
      }
      if (glane) {
        double* gi = g + (i0 + ns) * b + lane;
        gi[0] += cgx; gi[D] += cgv;
      }
      // ... and, with the first pass, the unary factor of the last state
      if (i0 == 0) {
        __syncwarp();
        if (hlane) Hd[(N - 1) * BD + dxx] += __ldg(rowM + (size_t)((N - 1) * CI) * RS);
        if (glane) g[(N - 1) * b + lane] += __ldg(rowG + (size_t)((N - 1) * CI) * RS);
      }
      __syncwarp();
    }
  }

  __device__ void assemble_obstacles_generic(const double* __restrict__ ml, int RS) {
    const int p = tp, q = tq;
    const int dxx = p * (p + 1) / 2 + q, dvx1 = (D + p) * (D + p + 1) / 2 + q, dvx2 = (D + q) * (D + q + 1) / 2 + p,
              dvv = (D + p) * (D + p + 1) / 2 + D + q;
    const int o1 = p * b + q, o2 = q * b + p;
    const bool offdiag = p != q, hlane = lane < T, glane = lane < D;
    const double* rowM = ml + (hlane ? lane : 0);          // entry `lane` of configuration c: rowM[c * RS]
    const double* rowG = ml + T + (glane ? lane : 0);
    for (int c0 = 0; c0 < C; c0 += 32) {
      int ri = c0 / (K + 1), rj = c0 - ri * (K + 1);   // (i, j) of the first configuration of the round
#pragma unroll 1
      for (int round = 0; round < 4; round++) {
        const int ci0 = c0 + round * 8;
        if (ci0 >= C) break;
        const int nt = min(8, C - ci0);
        int t = 0;
#pragma unroll 1
        while (t < nt) {
          // one segment = the configurations of interval ri present in this round
          double a0xx = 0, a0xv = 0, a0vv = 0, a1xx = 0, a1xv = 0, a1vv = 0, oxx = 0, oxv = 0, ovx = 0, ovv = 0;
          double g0x = 0, g0v = 0, g1x = 0, g1v = 0;
          const double* sm = rowM + (size_t)(ci0 + t) * RS;
          const double* sg = rowG + (size_t)(ci0 + t) * RS;
          if (rj == 0) {   // unary factor on x_i
            if (hlane) a0xx = __ldg(sm);
            if (glane) g0x = __ldg(sg);
            sm += RS; sg += RS; t++; rj = 1;
          }
          const int jn = min(K + 1 - rj, nt - t);   // GP obstacle factors of this segment
#pragma unroll 1
          for (int u = 0; u < jn; u++) {
            const double val = hlane ? __ldg(sm + u * RS) : 0.0;
            const double gval = glane ? __ldg(sg + u * RS) : 0.0;
            const double* ww = st.gpww[rj - 1 + u];
            const double* w = st.gpw[rj - 1 + u];
            a0xx = fma(ww[0], val, a0xx); a0xv = fma(ww[1], val, a0xv); a0vv = fma(ww[2], val, a0vv);
            oxx = fma(ww[3], val, oxx);   oxv = fma(ww[4], val, oxv);
            ovx = fma(ww[5], val, ovx);   ovv = fma(ww[6], val, ovv);
            a1xx = fma(ww[7], val, a1xx); a1xv = fma(ww[8], val, a1xv); a1vv = fma(ww[9], val, a1vv);
            g0x = fma(w[0], gval, g0x); g0v = fma(w[1], gval, g0v);
            g1x = fma(w[2], gval, g1x); g1v = fma(w[3], gval, g1v);
          }
          t += jn; rj += jn;
          if (hlane) {
            double* Hdi = Hd + ri * BD;
            const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
            if (jn > 0) {
              double* Hoi = Ho + ri * BB;
              double* Hdn = Hdi + BD;
              const double q0 = Hoi[o1], q1 = Hoi[o1 + D], q2 = Hoi[o1 + D * b], q3 = Hoi[o1 + D * b + D];
              const double q4 = Hoi[o2], q5 = Hoi[o2 + D], q6 = Hoi[o2 + D * b], q7 = Hoi[o2 + D * b + D];
              const double n0 = Hdn[dxx], n1 = Hdn[dvx1], n2 = Hdn[dvx2], n3 = Hdn[dvv];
              Hoi[o1] = q0 + oxx;               Hoi[o1 + D] = q1 + oxv;
              Hoi[o1 + D * b] = q2 + ovx;       Hoi[o1 + D * b + D] = q3 + ovv;
              if (offdiag) {
                Hoi[o2] = q4 + oxx;             Hoi[o2 + D] = q5 + oxv;
                Hoi[o2 + D * b] = q6 + ovx;     Hoi[o2 + D * b + D] = q7 + ovv;
              }
              Hdn[dxx] = n0 + a1xx;
              Hdn[dvx1] = n1 + a1xv;
              if (offdiag) Hdn[dvx2] = n2 + a1xv;
              Hdn[dvv] = n3 + a1vv;
            }
            Hdi[dxx] = h0 + a0xx;
            Hdi[dvx1] = h1 + a0xv;
            if (offdiag) Hdi[dvx2] = h2 + a0xv;
            Hdi[dvv] = h3 + a0vv;
          }
          if (glane) {
            double* gi = g + ri * b + lane;
            const double t0 = gi[0], t1 = gi[D];
            if (jn > 0) {
              const double t2 = gi[b], t3 = gi[b + D];
              gi[b] = t2 + g1x;
              gi[b + D] = t3 + g1v;
            }
            gi[0] = t0 + g0x;
            gi[D] = t1 + g0v;
          }
          if (rj > K) { rj = 0; ri++; }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  }

  // ---- (H + lambda I) delta = -g by a TWO-SIDED ("twisted") block-tridiagonal Cholesky: blocks 0..m-1 are
  //      eliminated downwards and blocks N-1..m+1 upwards in the same instruction stream, meeting at the
  //      middle block m = N/2 -- half the dependent chain of a one-way sweep.
  //      A panel = rows of [D_i ; rhs_i^T ; coupling block] lives in registers, one row per lane:
  //        lanes 0..b-1   row r of D_i (lower part)          -> row r of L_ii (diagonal stored as 1/l_rr)
  //        lane  b        rhs_i^T                            -> y_i^T (forward-substituted)
  //        lanes 16..16+b-1  row r of the coupling block     -> row r of X = H_{i+1,i} L_ii^-T (top sweep)
  //                                                             or Y = H_{i-1,i} L_ii^-T (bottom sweep)
  //      Every lane carries one row of the top panel (a[]) and one of the bottom panel (e[]).
  //      In place: Hd[i] <- L_ii, Ho[i] <- X_i (i < m) or Y_{i+1} (i >= m), both row-major [r][k].
  //      Returns false on a non-positive pivot (GTSAM's IndeterminantLinearSystemException). ----
  static __device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RSQ64H, ~2^-22 relative
#if GPMP2B_RSQRT_HALLEY
    // one third-order step: y (1 + r/2 + 3 r^2/8), r = 1 - x y^2 ~ 2^-21  ->  relative error 5/16 r^3 ~ 2^-65
    const double r = fma(-x * y, y, 1.0);
    return fma(y * r, fma(0.375, r, 0.5), y);
#else
    double r = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, r, y);
    r = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, r, y);
    return y;
#endif
  }

  // Right-looking Cholesky of the two panels held one row per lane (a: top panel, e: bottom panel).
  // The column loop is ROLLED with a sliding register window: a[c'] always holds the row's entry in column
  // k + c', so every register index is static and the ~100-instruction body stays resident in the L0
  // instruction cache (the kernel is instruction-fetch bound: an unrolled 14-column version is 10 KB of
  // straight-line code per panel).  Finished entries go straight to their final place (stT / stB point at this
  // lane's row, contiguous in the column index for every role).  The next pivot is computed one step AHEAD:
  // lane k+1 knows its own l_{k+1,k}, hence its updated diagonal, before the column is broadcast -- the shuffle +
  // rsqrt chain of column k+1 overlaps the shared-memory broadcast and update of column k.
  __device__ __forceinline__ void panel_factor(double (&a)[b], double (&e)[b], double* stT, double* stB, int nvT, int nvB) {
    double iT = fast_rsqrt(__shfl_sync(FULL_MASK, a[0], 0)), iB = fast_rsqrt(__shfl_sync(FULL_MASK, e[0], 0));
    // two rolled loops: the first b/2 columns update the full (b-1)-wide window, the remaining ones only the
    // b/2 - 1 entries that are still inside the matrix -- three quarters of the multiply-adds of one full-width loop
    panel_columns<b - 1>(a, e, stT, stB, nvT, nvB, iT, iB, 0, b / 2);
    panel_columns<b / 2 - 1>(a, e, stT, stB, nvT, nvB, iT, iB, b / 2, b);
  }
  // The damping lambda is added to the diagonals in shared memory before the sweep and a non-positive pivot is
  // not tested here: its rsqrt is NaN / inf, which reaches the solution vector, tested once at the end of solve().
  // Column k is broadcast through colbuf[k & 1] (double buffer); for even k it is stored one slot up so that the
  // read base (row k + 1) is always 16-byte aligned and the broadcast reads are LDS.128.
  template <int W>
  __device__ __forceinline__ void panel_columns(double (&a)[b], double (&e)[b], double* stT, double* stB, int nvT,
                                                int nvB, double& iT, double& iB, int k0, int k1) {
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
      const double lT = a[0] * iT, lB = e[0] * iB;
      if (k < nvT) stT[k] = (lane == k) ? iT : lT;     // diagonal stored as 1/l_kk
      if (k < nvB) stB[k] = (lane == k) ? iB : lB;
      const int sh = (k & 1) ^ 1;
      double* cb = colbuf + (k & 1) * 72 + sh;
      cb[lane] = lT;
      cb[36 + lane] = lB;
      // look-ahead: pivot of column k+1 (valid in lane k+1)
      const double pT = __shfl_sync(FULL_MASK, fma(-lT, lT, a[1]), (k + 1) & 31);
      const double pB = __shfl_sync(FULL_MASK, fma(-lB, lB, e[1]), (k + 1) & 31);
      iT = fast_rsqrt(pT);
      iB = fast_rsqrt(pB);
      __syncwarp();
      const double2* cT = reinterpret_cast<const double2*>(cb + k + 1);
      const double2* cB = reinterpret_cast<const double2*>(cb + 36 + k + 1);
#pragma unroll
      for (int c2 = 0; c2 < (W + 1) / 2; c2++) {
        const double2 t = cT[c2], u = cB[c2];
        a[2 * c2] = fma(-lT, t.x, a[2 * c2 + 1]);
        e[2 * c2] = fma(-lB, u.x, e[2 * c2 + 1]);
        if (2 * c2 + 1 < W) {
          a[2 * c2 + 1] = fma(-lT, t.y, a[2 * c2 + 2]);
          e[2 * c2 + 1] = fma(-lB, u.y, e[2 * c2 + 2]);
        }
      }
    }
  }

  // D_tgt -= Z Z^T, Z row-major b x b: this lane's run of <= 4 entries of row sch_r (its row of Z stays in registers)
  __device__ __forceinline__ void schur(const double* Z, double* Dtgt) {
    double2 zr[b / 2];
    {
      const double2* zp = reinterpret_cast<const double2*>(Z + sch_r * b);
#pragma unroll
      for (int k2 = 0; k2 < b / 2; k2++) zr[k2] = zp[k2];
    }
    double* dt = Dtgt + sch_r * (sch_r + 1) / 2 + sch_c0;
#if GPMP2B_SCHUR_ROTATE
    // rows of Z that are 8 apart share their four banks (b = 14: 112-byte rows): the lanes whose run starts at column 8
    // or 12 walk it rotated (sch_rot), so that in one quarter-warp they never read row rho + 8 while another reads rho
    const double2* zc0 = reinterpret_cast<const double2*>(Z + sch_c0 * b);
    int jj = sch_rot;
#pragma unroll 1
    for (int j = 0; j < sch_n; j++) {
      const double2* zc = zc0 + jj * (b / 2);
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int k2 = 0; k2 < b / 2; k2++) {
        const double2 v = zc[k2];
        acc0 = fma(zr[k2].x, v.x, acc0);
        acc1 = fma(zr[k2].y, v.y, acc1);
      }
      dt[jj] -= acc0 + acc1;
      jj = (jj + 1 == sch_n) ? 0 : jj + 1;
    }
#else
    const double2* zc = reinterpret_cast<const double2*>(Z + sch_c0 * b);
#pragma unroll 1
    for (int j = 0; j < sch_n; j++) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int k2 = 0; k2 < b / 2; k2++) {
        const double2 v = zc[k2];
        acc0 = fma(zr[k2].x, v.x, acc0);
        acc1 = fma(zr[k2].y, v.y, acc1);
      }
      dt[j] -= acc0 + acc1;
      zc += b / 2;
    }
#endif
  }
  // rhs_tgt[r] -= sum_k Z[r][k] y[k]   (lanes r < b)
  __device__ __forceinline__ void rhs_update(const double* Z, const double* y, double* tgt) {
    if (lane < b) {
      const double2* zr = reinterpret_cast<const double2*>(Z + lane * b);
      const double2* yy = reinterpret_cast<const double2*>(y);
      double acc = 0.0;
#pragma unroll
      for (int k2 = 0; k2 < b / 2; k2++) {
        const double2 u = zr[k2], v = yy[k2];
        acc = fma(u.x, v.x, acc);
        acc = fma(u.y, v.y, acc);
      }
      tgt[lane] -= acc;
    }
  }

  // =====================================================================================================
  // STREAMED variant of assemble + solve (phase-kernel pipeline, layout 3): the coupling blocks H_{i,i+1} -- 63 % of
  // the H storage -- never sit in shared memory as a whole.  Each is assembled (template + obstacle sums from the
  // M-list) into a one-block window right before the sweep that consumes it, its factor X_i is kept in the window for
  // the Schur / rhs updates of that step and written TRANSPOSED to a per-block global slab (L2-resident), from which
  // the back substitution reads its column contiguously.  17 KB instead of 32 KB per trajectory: 12 solves per SM
  // instead of 7.  Same arithmetic, same association order as assemble_from_global() + solve(): bit-identical.
  // =====================================================================================================
  struct EntryIdx {
    int dxx, dvx1, dvx2, dvv, o1, o2;
    bool offdiag, hlane, glane;
  };
  __device__ __forceinline__ EntryIdx entry_idx() const {
    EntryIdx e;
    const int p = tp, q = tq;
    e.dxx = p * (p + 1) / 2 + q; e.dvx1 = (D + p) * (D + p + 1) / 2 + q; e.dvx2 = (D + q) * (D + q + 1) / 2 + p;
    e.dvv = (D + p) * (D + p + 1) / 2 + D + q;
    e.o1 = p * b + q; e.o2 = q * b + p;
    e.offdiag = p != q; e.hlane = lane < T; e.glane = lane < D;
    return e;
  }

  // diagonal blocks + gradient only (the a0 / a1 / g sums of assemble_obstacles_aligned, Ho left out)
  template <int KS>
  __device__ void assemble_diag_aligned(const double* __restrict__ ml, int RS) {
    constexpr int CI = KS + 1, IPP = 32 / CI;
    const EntryIdx E = entry_idx();
    const bool offdiag = E.offdiag, hlane = E.hlane, glane = E.glane;
    const int dxx = E.dxx, dvx1 = E.dvx1, dvx2 = E.dvx2, dvv = E.dvv;
    const double* rowM = ml + (hlane ? lane : 0);          // entry `lane` of configuration c: rowM[c * RS]
    const double* rowG = ml + T + (glane ? lane : 0);
#pragma unroll 1
    for (int i0 = 0; i0 < N - 1; i0 += IPP) {
      double cxx = 0.0, cxv = 0.0, cvv = 0.0, cgx = 0.0, cgv = 0.0;
      const int ns = min(IPP, N - 1 - i0);
#pragma unroll 1
      for (int sl = 0; sl < ns; sl++) {
        const int i = i0 + sl;
        const double* vm = rowM + (size_t)(i * CI) * RS;
        const double* vg = rowG + (size_t)(i * CI) * RS;
        double val[CI], gval[CI];
#pragma unroll
        for (int j = 0; j < CI; j++) { val[j] = hlane ? __ldg(vm + j * RS) : 0.0; gval[j] = glane ? __ldg(vg + j * RS) : 0.0; }
        double a0xx = cxx + val[0], a0xv = cxv, a0vv = cvv, a1xx = 0, a1xv = 0, a1vv = 0;
        double g0x = cgx + gval[0], g0v = cgv, g1x = 0, g1v = 0;
#pragma unroll
        for (int j = 1; j <= KS; j++) {
          a0xx = fma(st.gpww[j - 1][0], val[j], a0xx); a0xv = fma(st.gpww[j - 1][1], val[j], a0xv); a0vv = fma(st.gpww[j - 1][2], val[j], a0vv);
          a1xx = fma(st.gpww[j - 1][7], val[j], a1xx); a1xv = fma(st.gpww[j - 1][8], val[j], a1xv); a1vv = fma(st.gpww[j - 1][9], val[j], a1vv);
          g0x = fma(st.gpw[j - 1][0], gval[j], g0x); g0v = fma(st.gpw[j - 1][1], gval[j], g0v);
          g1x = fma(st.gpw[j - 1][2], gval[j], g1x); g1v = fma(st.gpw[j - 1][3], gval[j], g1v);
        }
        if (hlane) {
          double* Hdi = Hd + i * BD;
          const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
          Hdi[dxx] = h0 + a0xx;
          Hdi[dvx1] = h1 + a0xv;
          if (offdiag) Hdi[dvx2] = h2 + a0xv;
          Hdi[dvv] = h3 + a0vv;
        }
        if (glane) {
          double* gi = g + i * b + lane;
          const double t0 = gi[0], t1 = gi[D];
          gi[0] = t0 + g0x;
          gi[D] = t1 + g0v;
        }
        cxx = a1xx; cxv = a1xv; cvv = a1vv; cgx = g1x; cgv = g1v;
        __syncwarp();
      }
      if (hlane) {
        double* Hdi = Hd + (i0 + ns) * BD;
        const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
        Hdi[dxx] = h0 + cxx; Hdi[dvx1] = h1 + cxv; if (offdiag) Hdi[dvx2] = h2 + cxv; Hdi[dvv] = h3 + cvv;
      }
      if (glane) {
        double* gi = g + (i0 + ns) * b + lane;
        gi[0] += cgx; gi[D] += cgv;
      }
      if (i0 == 0) {
        __syncwarp();
        if (hlane) Hd[(N - 1) * BD + dxx] += __ldg(rowM + (size_t)((N - 1) * CI) * RS);
        if (glane) g[(N - 1) * b + lane] += __ldg(rowG + (size_t)((N - 1) * CI) * RS);
      }
      __syncwarp();
    }
  }
  // generic obs_check_inter: the rounds / segments of assemble_obstacles_generic, diagonal blocks and gradient only
  __device__ void assemble_diag_generic(const double* __restrict__ ml, int RS) {
    const EntryIdx E = entry_idx();
    const bool offdiag = E.offdiag, hlane = E.hlane, glane = E.glane;
    const int dxx = E.dxx, dvx1 = E.dvx1, dvx2 = E.dvx2, dvv = E.dvv;
    const double* rowM = ml + (hlane ? lane : 0);          // entry `lane` of configuration c: rowM[c * RS]
    const double* rowG = ml + T + (glane ? lane : 0);
    for (int c0 = 0; c0 < C; c0 += 32) {
      int ri = c0 / (K + 1), rj = c0 - ri * (K + 1);
#pragma unroll 1
      for (int round = 0; round < 4; round++) {
        const int ci0 = c0 + round * 8;
        if (ci0 >= C) break;
        const int nt = min(8, C - ci0);
        int t = 0;
#pragma unroll 1
        while (t < nt) {
          double a0xx = 0, a0xv = 0, a0vv = 0, a1xx = 0, a1xv = 0, a1vv = 0;
          double g0x = 0, g0v = 0, g1x = 0, g1v = 0;
          const double* sm = rowM + (size_t)(ci0 + t) * RS;
          const double* sg = rowG + (size_t)(ci0 + t) * RS;
          if (rj == 0) {
            if (hlane) a0xx = __ldg(sm);
            if (glane) g0x = __ldg(sg);
            sm += RS; sg += RS; t++; rj = 1;
          }
          const int jn = min(K + 1 - rj, nt - t);
#pragma unroll 1
          for (int u = 0; u < jn; u++) {
            const double val = hlane ? __ldg(sm + u * RS) : 0.0;
            const double gval = glane ? __ldg(sg + u * RS) : 0.0;
            const double* ww = st.gpww[rj - 1 + u];
            const double* w = st.gpw[rj - 1 + u];
            a0xx = fma(ww[0], val, a0xx); a0xv = fma(ww[1], val, a0xv); a0vv = fma(ww[2], val, a0vv);
            a1xx = fma(ww[7], val, a1xx); a1xv = fma(ww[8], val, a1xv); a1vv = fma(ww[9], val, a1vv);
            g0x = fma(w[0], gval, g0x); g0v = fma(w[1], gval, g0v);
            g1x = fma(w[2], gval, g1x); g1v = fma(w[3], gval, g1v);
          }
          t += jn; rj += jn;
          if (hlane) {
            double* Hdi = Hd + ri * BD;
            const double h0 = Hdi[dxx], h1 = Hdi[dvx1], h2 = Hdi[dvx2], h3 = Hdi[dvv];
            if (jn > 0) {
              double* Hdn = Hdi + BD;
              const double n0 = Hdn[dxx], n1 = Hdn[dvx1], n2 = Hdn[dvx2], n3 = Hdn[dvv];
              Hdn[dxx] = n0 + a1xx;
              Hdn[dvx1] = n1 + a1xv;
              if (offdiag) Hdn[dvx2] = n2 + a1xv;
              Hdn[dvv] = n3 + a1vv;
            }
            Hdi[dxx] = h0 + a0xx;
            Hdi[dvx1] = h1 + a0xv;
            if (offdiag) Hdi[dvx2] = h2 + a0xv;
            Hdi[dvv] = h3 + a0vv;
          }
          if (glane) {
            double* gi = g + ri * b + lane;
            const double t0 = gi[0], t1 = gi[D];
            if (jn > 0) {
              const double t2 = gi[b], t3 = gi[b + D];
              gi[b] = t2 + g1x;
              gi[b + D] = t3 + g1v;
            }
            gi[0] = t0 + g0x;
            gi[D] = t1 + g0v;
          }
          if (rj > K) { rj = 0; ri++; }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  }
  // H_{i,i+1} = template + obstacle sums of interval i, into the window Zw (row-major b x b).  The sums are
  // associated as the fused kernels do: aligned K in one piece; generic K in the pieces the 8-configuration rounds
  // cut an interval into (each piece summed from zero, then added to the block).
  __device__ __forceinline__ void assemble_off_block(const double* __restrict__ ml, int RS, int i, double* Zw) const {
    const EntryIdx E = entry_idx();
    if (!E.hlane) return;
    const double* rowM = ml + lane;
    const double* __restrict__ hc = hconst + i * BB;
    const int o1 = E.o1, o2 = E.o2;
    double b0 = __ldg(hc + o1), b1 = __ldg(hc + o1 + D), b2 = __ldg(hc + o1 + D * b), b3 = __ldg(hc + o1 + D * b + D);
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (E.offdiag) { m0 = __ldg(hc + o2); m1 = __ldg(hc + o2 + D); m2 = __ldg(hc + o2 + D * b); m3 = __ldg(hc + o2 + D * b + D); }
    const int c_lo = i * (K + 1) + 1, c_hi = (i + 1) * (K + 1);     // the GP-interpolated configurations of interval i
    const bool aligned = GPMP2B_ALIGNED_ACC && (K == 5 || K == 4);
    int c = c_lo;
#pragma unroll 1
    while (c < c_hi) {
      const int ce = aligned ? c_hi : min(c_hi, (c & ~7) + 8);
      double oxx = 0, oxv = 0, ovx = 0, ovv = 0;
#pragma unroll 1
      for (; c < ce; c++) {
        const double val = __ldg(rowM + (size_t)c * RS);
        const double* ww = st.gpww[c - c_lo];
        oxx = fma(ww[3], val, oxx); oxv = fma(ww[4], val, oxv);
        ovx = fma(ww[5], val, ovx); ovv = fma(ww[6], val, ovv);
      }
      b0 += oxx; b1 += oxv; b2 += ovx; b3 += ovv;
      if (E.offdiag) { m0 += oxx; m1 += oxv; m2 += ovx; m3 += ovv; }
    }
    Zw[o1] = b0; Zw[o1 + D] = b1; Zw[o1 + D * b] = b2; Zw[o1 + D * b + D] = b3;
    if (E.offdiag) { Zw[o2] = m0; Zw[o2 + D] = m1; Zw[o2 + D * b] = m2; Zw[o2 + D * b + D] = m3; }
  }

  // everything that lands in Hd and g (template, per-state pass, obstacle sums); the coupling blocks follow in the sweep
  __device__ void assemble_diag_from_global(const double* __restrict__ ml, int RS) {
    copy_in(reinterpret_cast<const double2*>(hconst + (N - 1) * BB), reinterpret_cast<double2*>(Hd), (N * BD + 1) / 2);
    __syncwarp();
    state_pass<false, true>();
    __syncwarp();
#if GPMP2B_ALIGNED_ACC
    if (K == 5) { assemble_diag_aligned<5>(ml, RS); return; }
    if (K == 4) { assemble_diag_aligned<4>(ml, RS); return; }
#endif
    assemble_diag_generic(ml, RS);
  }

  // solve() with the coupling blocks streamed (see above).  slab: (N-1) * BB doubles of this block's global scratch.
  __device__ bool solve_streamed(double lambda, const double* __restrict__ ml, int RS, double* slab) {
    static_assert(b <= 15 && (b % 2) == 0, "panel layout needs b + 1 <= 16 lanes per half-warp");
    for (int idx = lane; idx < N * b; idx += 32) {
      dl[idx] = -g[idx];
      if (lambda != 0.0) {
        const int blk = idx / b, rr = idx - blk * b;
        Hd[blk * BD + rr * (rr + 1) / 2 + rr] += lambda;
      }
    }
    const int r = lane & 15;
    const bool isD = lane < b, isR = lane == b, isO = lane >= 16 && r < b;
    const int m = N / 2;
    double *ldT = Hd, *stT = Hd, *ldB = Hd;
    int ldsT = 1, inc = 0, nv = 0;
    if (isD) { ldT = stT = Hd + r * (r + 1) / 2; ldB = Hd + (N - 1) * BD + r * (r + 1) / 2; inc = BD; nv = r + 1; }
    else if (isR) { ldT = stT = dl; ldB = dl + (N - 1) * b; inc = b; nv = b; }
    else if (isO) { ldT = ZT + r; ldsT = b; stT = ZT + r * b; ldB = ZB + r * b; inc = 0; nv = b; }   // the windows do not move
    __syncwarp();

    double a[b], e[b];
#pragma unroll 1
    for (int t = 0; t <= m; t++) {
      const int iT = t, iB = N - 1 - t;
      const bool mid = t == m, haveB = !mid && iB > m;
      // this step's coupling blocks into the windows: H_{t,t+1} for the top sweep, H_{iB-1,iB} for the bottom sweep
      if (!mid) {
        assemble_off_block(ml, RS, iT, ZT);
        if (haveB) assemble_off_block(ml, RS, iB - 1, ZB);
        __syncwarp();
      }
      const int nvT = (mid && isO) ? 0 : nv, nvB = haveB ? nv : 0;
#pragma unroll
      for (int c = 0; c < b; c++) {
        a[c] = (c < nvT) ? ldT[c * ldsT] : 0.0;
        e[c] = (c < nvB) ? ldB[c] : ((isD && c == r) ? 1.0 : 0.0);
      }
      __syncwarp();            // (the factor rows overwrite the windows in place: every lane has loaded its panel row)
      panel_factor(a, e, stT, ldB, nvT, nvB);
      ldT += inc; stT += inc; ldB -= inc;
      __syncwarp();
      if (!mid) {
#pragma unroll 1
        for (int side = 0; side < (haveB ? 2 : 1); side++) {
          const int tg = side ? iB - 1 : iT + 1, yb = side ? iB : iT;
          const double* Z = side ? ZB : ZT;
          schur(Z, Hd + tg * BD);
          rhs_update(Z, dl + yb * b, dl + tg * b);
          // keep the factor for the back substitution, transposed: slab[blk][k][r] = Z[r][k]
          double* dst = slab + (side ? iB - 1 : iT) * BB;
          for (int d = lane; d < BB; d += 32) {
            const int k = d / b, rr = d - k * b;
            dst[d] = Z[rr * b + k];
          }
        }
        __syncwarp();
      }
    }
    // back substitution from the middle outwards (as solve(); the coupling factors come from the slab)
    const int hw = lane >> 4;
#pragma unroll 1
    for (int t = m; t >= 0; t--) {
      const int iT = t, iB = N - 1 - t;
      const bool mid = t == m, haveB = !mid && iB > m;
      const int blk = (hw && !mid) ? iB : iT;
      const bool act = r < b && (hw == 0 || haveB);
      const double* Lb = Hd + blk * BD;
      double tt = act ? dl[blk * b + r] : 0.0;
      double Lc[b];
#pragma unroll
      for (int k = 0; k < b; k++) Lc[k] = (act && k > r) ? Lb[k * (k + 1) / 2 + r] : 0.0;
      const double dr = act ? Lb[r * (r + 1) / 2 + r] : 0.0;
      if (act && !mid) {
        const double2* Zt = reinterpret_cast<const double2*>(slab + (hw ? iB - 1 : iT) * BB + r * b);   // column r of Z
        const double* xn = dl + (hw ? iB - 1 : iT + 1) * b;
        double t2 = 0.0;
#pragma unroll
        for (int rr = 0; rr < b; rr += 2) {
          const double2 z = __ldcg(Zt + rr / 2);
          tt = fma(-z.x, xn[rr], tt);
          t2 = fma(-z.y, xn[rr + 1], t2);
        }
        tt += t2;
      }
#pragma unroll
      for (int k = b - 1; k >= 0; k--) {
        const double xk = __shfl_sync(FULL_MASK, tt * dr, k, 16);
        tt = fma(-Lc[k], xk, tt);
      }
      tt *= dr;
      if (act) dl[blk * b + r] = tt;
      __syncwarp();
    }
    bool ok = true;
    for (int idx = lane; idx < N * b; idx += 32) ok = ok && (fabs(dl[idx]) < CUDART_INF);
    return __all_sync(FULL_MASK, ok);
  }

  __device__ bool solve(double lambda) {
    static_assert(b <= 15 && (b % 2) == 0, "panel layout needs b + 1 <= 16 lanes per half-warp");
    for (int idx = lane; idx < N * b; idx += 32) {
      dl[idx] = -g[idx];
      if (lambda != 0.0) {                       // damping: H + lambda I (the Schur updates are linear in the diagonal)
        const int blk = idx / b, rr = idx - blk * b;
        Hd[blk * BD + rr * (rr + 1) / 2 + rr] += lambda;
      }
    }
    const int r = lane & 15;
    const bool isD = lane < b, isR = lane == b, isO = lane >= 16 && r < b;
    const int m = N / 2;   // middle block; top sweep 0..m-1, bottom sweep N-1..m+1
    // per-lane row pointers: load (top may be strided), store (always contiguous), per-step increments
    double *ldT = Hd, *stT = Hd, *ldB = Hd;
    int ldsT = 1, inc = 0, nv = 0;
    if (isD) { ldT = stT = Hd + r * (r + 1) / 2; ldB = Hd + (N - 1) * BD + r * (r + 1) / 2; inc = BD; nv = r + 1; }
    else if (isR) { ldT = stT = dl; ldB = dl + (N - 1) * b; inc = b; nv = b; }
    else if (isO) { ldT = Ho + r; ldsT = b; stT = Ho + r * b; ldB = Ho + (N - 2) * BB + r * b; inc = BB; nv = b; }
    __syncwarp();

    // forward sweeps; the last step (t == m) is the middle block: top panel only, no coupling rows
    double a[b], e[b];
#pragma unroll 1
    for (int t = 0; t <= m; t++) {
      const int iT = t, iB = N - 1 - t;
      const bool mid = t == m, haveB = !mid && iB > m;
      const int nvT = (mid && isO) ? 0 : nv, nvB = haveB ? nv : 0;
#pragma unroll
      for (int c = 0; c < b; c++) {
        a[c] = (c < nvT) ? ldT[c * ldsT] : 0.0;
        e[c] = (c < nvB) ? ldB[c] : ((isD && c == r) ? 1.0 : 0.0);
      }
      panel_factor(a, e, stT, ldB, nvT, nvB);
      ldT += inc; stT += inc; ldB -= inc;
      __syncwarp();
      if (!mid) {
        // Schur complements into the neighbouring blocks (both may be the middle block: same lane, in order)
#pragma unroll 1
        for (int side = 0; side < (haveB ? 2 : 1); side++) {
          const int zb = side ? iB - 1 : iT, tg = side ? iB - 1 : iT + 1, yb = side ? iB : iT;
          schur(Ho + zb * BB, Hd + tg * BD);
          rhs_update(Ho + zb * BB, dl + yb * b, dl + tg * b);
        }
        __syncwarp();
      }
    }
    // back substitution from the middle outwards: lower half-warp upwards (x_i = L_ii^-T (y_i - X_i^T x_{i+1})),
    // upper half-warp downwards (x_i = L_ii^-T (y_i - Y_i^T x_{i-1}))
    const int hw = lane >> 4;
#pragma unroll 1
    for (int t = m; t >= 0; t--) {
      const int iT = t, iB = N - 1 - t;
      const bool mid = t == m, haveB = !mid && iB > m;
      const int blk = (hw && !mid) ? iB : iT;
      const bool act = r < b && (hw == 0 || haveB);
      const double* Lb = Hd + blk * BD;
      double tt = act ? dl[blk * b + r] : 0.0;
#if GPMP2B_BACKSUB_REG
      // lane r keeps column r of L_blk below the diagonal (zeros elsewhere) and 1/l_rr in registers: the 14 loads
      // are in flight together and the dependent chain of a step is DMUL -> SHFL -> DFMA (it was LDS -> DMUL ->
      // SHFL -> LDS -> DFMA behind a rolled loop with divergent branches: ~250 cycles per step on the profile)
      double Lc[b];
#pragma unroll
      for (int k = 0; k < b; k++) Lc[k] = (act && k > r) ? Lb[k * (k + 1) / 2 + r] : 0.0;
      const double dr = act ? Lb[r * (r + 1) / 2 + r] : 0.0;
      if (act && !mid) {
        const double* Z = Ho + (hw ? iB - 1 : iT) * BB;
        const double* xn = dl + (hw ? iB - 1 : iT + 1) * b;
        double t2 = 0.0;
#pragma unroll
        for (int rr = 0; rr < b; rr += 2) {
          tt = fma(-Z[rr * b + r], xn[rr], tt);
          t2 = fma(-Z[(rr + 1) * b + r], xn[rr + 1], t2);
        }
        tt += t2;
      }
#pragma unroll
      for (int k = b - 1; k >= 0; k--) {
        const double xk = __shfl_sync(FULL_MASK, tt * dr, k, 16);
        tt = fma(-Lc[k], xk, tt);      // no-op for lanes r >= k (Lc[k] = 0)
      }
      tt *= dr;                        // lane r's t_r is final after step r + 1
#else
      if (act && !mid) {
        const double* Z = Ho + (hw ? iB - 1 : iT) * BB;
        const double* xn = dl + (hw ? iB - 1 : iT + 1) * b;
#pragma unroll
        for (int rr = 0; rr < b; rr++) tt = fma(-Z[rr * b + r], xn[rr], tt);
      }
#pragma unroll 1
      for (int k = b - 1; k >= 0; k--) {
        const int kd = k * (k + 1) / 2;
        double xk = (act && r == k) ? tt * Lb[kd + k] : 0.0;
        xk = __shfl_sync(FULL_MASK, xk, k, 16);
        if (r == k) tt = xk;
        else if (act && r < k) tt = fma(-Lb[kd + r], xk, tt);
      }
#endif
      if (act) dl[blk * b + r] = tt;
      __syncwarp();
    }
    // a non-positive pivot (GTSAM: IndeterminantLinearSystemException) shows up as a non-finite solution
    bool ok = true;
    for (int idx = lane; idx < N * b; idx += 32) ok = ok && (fabs(dl[idx]) < CUDART_INF);
    return __all_sync(FULL_MASK, ok);
  }

  // g^T H g with the block-tridiagonal H still un-factored (Dogleg's steepest-descent step length,
  // GaussianBayesTree::optimizeGradientSearch: step = -g.g / |R g|^2)  -- warp-collective
  __device__ double g_dot_Hg() const {
    double acc = 0.0;
    for (int i = 0; i < N; i++) {
      if (lane < b) {
        const int r = lane;
        double hg = 0.0;
        const double* Hdi = Hd + i * BD;
        for (int c = 0; c < b; c++) {
          const int hi = r > c ? r : c, lo = r > c ? c : r;
          hg = fma(Hdi[hi * (hi + 1) / 2 + lo], g[i * b + c], hg);
        }
        if (i < N - 1) for (int c = 0; c < b; c++) hg = fma(Ho[i * BB + r * b + c], g[(i + 1) * b + c], hg);
        if (i > 0) for (int c = 0; c < b; c++) hg = fma(Ho[(i - 1) * BB + c * b + r], g[(i - 1) * b + c], hg);
        acc = fma(g[i * b + r], hg, acc);
      }
    }
    return warp_sum(acc);
  }

  // Values::retract for vector states: x + delta
  __device__ void accept_step() {
    for (int idx = lane; idx < N * b; idx += 32) xs[idx] += dl[idx];
    __syncwarp();
  }
  __device__ void step_back() {   // last_values of gpmp2::optimize (Gauss-Newton only)
    for (int idx = lane; idx < N * b; idx += 32) xs[idx] -= dl[idx];
    __syncwarp();
  }
  __device__ void debug_obs(int cidx, double* de, double* dc) {
    const int i = cidx / (K + 1), j = cidx - i * (K + 1);
    double e2 = 0.0, es = 0.0;
    config_error<D, NDIM, 0, true>(rb, sdf, config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, e2, es, de, dc);
  }

  __device__ void backup_H(double* dst) const {
    const int n2 = (N * BD + (N - 1) * BB + 1) / 2;   // Ho and Hd are contiguous in shared memory (Ho first)
    const double2* src = reinterpret_cast<const double2*>(Ho);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (int idx = lane; idx < n2; idx += 32) d2[idx] = src[idx];
  }
  // The backup slab is WRITTEN by this kernel (backup_H, plain stores) on every LM iteration, so it must not be read
  // through the non-coherent path (ld.global.nc requires memory that is read-only for the kernel's lifetime): the
  // restore uses ld.global.cg (L2-coherent) loads.  Only the H template -- written by a memcpy before launch -- is __ldg'd.
  __device__ void restore_H(const double* src) {
    copy_in<true>(reinterpret_cast<const double2*>(src), reinterpret_cast<double2*>(Ho), (N * BD + (N - 1) * BB + 1) / 2);
    __syncwarp();
  }
  // global -> shared copy with 16 loads in flight per lane (a plain loop exposes one L2 round trip per iteration:
  // measured ~25k cycles for the 25 KB H template on a lone warp)
  template <bool COHERENT = false>
  __device__ __forceinline__ void copy_in(const double2* src, double2* dst, int n2) const {
    auto ld = [](const double2* p) { return COHERENT ? __ldcg(p) : __ldg(p); };
#pragma unroll 1
    for (int base = lane; base < n2; base += 32 * 8) {
      double2 r0 = ld(src + min(base, n2 - 1)), r1 = ld(src + min(base + 32, n2 - 1));
      double2 r2 = ld(src + min(base + 64, n2 - 1)), r3 = ld(src + min(base + 96, n2 - 1));
      double2 r4 = ld(src + min(base + 128, n2 - 1)), r5 = ld(src + min(base + 160, n2 - 1));
      double2 r6 = ld(src + min(base + 192, n2 - 1)), r7 = ld(src + min(base + 224, n2 - 1));
      dst[base] = r0;
      if (base + 32 < n2) dst[base + 32] = r1;
      if (base + 64 < n2) dst[base + 64] = r2;
      if (base + 96 < n2) dst[base + 96] = r3;
      if (base + 128 < n2) dst[base + 128] = r4;
      if (base + 160 < n2) dst[base + 160] = r5;
      if (base + 192 < n2) dst[base + 192] = r6;
      if (base + 224 < n2) dst[base + 224] = r7;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// The kernel.  grid = resident warps (persistent, grid-stride over problems), block = 32 threads.
// ------------------------------------------------------------------------------------------------
template <class Opt, int OPT>
__global__ void __launch_bounds__(32)
gpmp2b_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf,
              const __grid_constant__ KSetting st, const __grid_constant__ KProblem pr,
              const double* __restrict__ hconst, int mode) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim;
  Opt o(rb, sdf, st, hconst, smem);
  const int lane = o.lane, N = o.N, b = Opt::b;
  const int TL = 2 * N * D;
  unsigned long long n_lin = 0, n_solve = 0, n_err = 0;
#ifdef GPMP2B_PHASE_TIMING
  long long t_lin = 0, t_solve = 0, t_err = 0, t_bk = 0, t0_;
#define PT_BEGIN() t0_ = clock64()
#define PT_END(acc) acc += clock64() - t0_
#else
#define PT_BEGIN()
#define PT_END(acc)
#endif

  // Persistent warps pull problems from a global counter: LM trial counts differ per trajectory, and with a static
  // grid-stride assignment the slowest warp kept its SM busy for ~4 % of the kernel after the others had finished.
  auto next_problem = [&]() -> int64_t {
    unsigned long long nx = 0;
    if (lane == 0) nx = atomicAdd(pr.queue, 1ull);
    return (int64_t)__shfl_sync(FULL_MASK, nx, 0) + gridDim.x;
  };
  for (int64_t prob = blockIdx.x; prob < pr.B; prob = next_problem()) {
    // ---- load the trajectory (wire layout [x_0..x_T | v_0..v_T]) ----
    const double* tin = pr.init_traj + prob * TL;
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      o.xs[i * b + d] = tin[idx];
      o.xs[i * b + D + d] = tin[N * D + idx];
    }
    o.start_conf = pr.start_conf ? pr.start_conf + prob * D : nullptr;
    o.start_vel = pr.start_vel ? pr.start_vel + prob * D : nullptr;
    o.end_conf = pr.end_conf ? pr.end_conf + prob * D : nullptr;
    o.end_vel = pr.end_vel ? pr.end_vel + prob * D : nullptr;
    o.goal_pos_p = pr.goal_pos_pp ? pr.goal_pos_pp + prob * 3 : st.goal_pos;
    o.goal_R_p = pr.goal_R_pp ? pr.goal_R_pp + prob * 9 : st.goal_R;
    o.orient_R_p = pr.orient_R_pp ? pr.orient_R_pp + prob * 9 : st.orient_R;
    o.fix_conf = pr.fix_conf_pp ? pr.fix_conf_pp + prob * D : nullptr;
    o.fix_vel = pr.fix_vel_pp ? pr.fix_vel_pp + prob * D : nullptr;
    __syncwarp();

    if constexpr (OPT < 0) {   // ======== auxiliary kernel: collision cost + parity/debug modes ========
    if (mode == KMODE_COLLISION_COST) {
      const double cc = o.collision_cost();
      if (lane == 0) pr.out_coll_cost[prob] = cc;
      __syncwarp();
      continue;
    }
    if (mode == KMODE_OBS_ERRORS) {
      const int S = rb.n_spheres, C = o.C;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int cidx = c0 + lane;
        if (cidx < C) {
          double* de = pr.out_obs_err + ((size_t)prob * C + cidx) * S;
          double* dc = pr.out_centers ? pr.out_centers + ((size_t)prob * C + cidx) * S * 3 : nullptr;
          o.debug_obs(cidx, de, dc);
        }
      }
      __syncwarp();
      continue;
    }
    if (mode == KMODE_LINEARIZE) {
      // the error first: its SDF gathers borrow the H storage as scratch (err_scratch)
      for (int idx = lane; idx < N * b; idx += 32) o.dl[idx] = 0.0;
      __syncwarp();
      const double err = o.template eval_error<true>();
      o.linearize();
      // expand to the debug layout
      if (pr.out_Hdiag)
        for (int idx = lane; idx < N * b * b; idx += 32) {
          const int i = idx / (b * b), rc = idx - i * b * b, r = rc / b, c = rc - r * b;
          const int hi = r > c ? r : c, lo = r > c ? c : r;
          pr.out_Hdiag[(size_t)prob * N * b * b + idx] = o.Hd[i * Opt::BD + hi * (hi + 1) / 2 + lo];
        }
      if (pr.out_Hoff)
        for (int idx = lane; idx < (N - 1) * b * b; idx += 32) pr.out_Hoff[(size_t)prob * (N - 1) * b * b + idx] = o.Ho[idx];
      if (pr.out_g)
        for (int idx = lane; idx < N * b; idx += 32) pr.out_g[(size_t)prob * N * b + idx] = o.g[idx];
      if (pr.out_error && lane == 0) pr.out_error[prob] = err;
      __syncwarp();
      continue;
    }
    } else {   // ======== optimizer kernel (one instantiation per optimizer type) ========

    // ---- gpmp2::optimize (BatchTrajOptimizer.cpp:212-308) over LM / GN [GTSAM semantics, SURVEY App. B] ----
    double* hb = pr.h_backup + (size_t)blockIdx.x * h_backup_size(D, N);
    constexpr bool is_lm = OPT == 1;
    double lambda = 100.0;                     // setlambdaInitial(100.0), BatchTrajOptimizer.cpp:226
    double Delta = 0.2;                        // setDeltaInitial(0.2), BatchTrajOptimizer.cpp:219-222
    const double lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3;
    const double absoluteErrorTol = 1e-5, errorTol = 0.0, relativeErrorTol = st.rel_thresh;
    int iterations = 0, status = 0;
    for (int idx = lane; idx < N * b; idx += 32) o.dl[idx] = 0.0;   // error at xs = error at xs + 0
    __syncwarp();
    double error = o.template eval_error<true>();
    n_err++;
    double currentError = error;
    bool step_back = false;   // GN only: return last_values (= xs - dl)
    if (currentError <= errorTol) status |= 32;
    else if (iterations >= st.max_iter) status |= 4;
    else {
      bool converged = false;
      int why = 0;
      do {
        currentError = error;
        PT_BEGIN();
        o.linearize();
        PT_END(t_lin);
        n_lin++;
        if constexpr (OPT == 2) {
          // DoglegOptimizer::iterate -> DoglegOptimizerImpl::Iterate(ONE_STEP_PER_ITERATION) [GTSAM, recalled]
          double gg = 0.0;
          for (int idx = lane; idx < N * b; idx += 32) gg = fma(o.g[idx], o.g[idx], gg);
          gg = warp_sum(gg);
          const double gHg = o.g_dot_Hg();
          n_solve++;
          const bool solved = o.solve(0.0);           // dx_n = -H^-1 g  (in dl)
          if (!solved) { status |= 16; break; }
          double* nv = o.Ho;                          // H storage is free after the solve: keep dx_n there
          o.err_scratch_off = (o.N * Opt::b + 1) & ~1;   // ... and keep the error pass's scratch off it
          double gn = 0.0, nn = 0.0;
          for (int idx = lane; idx < N * b; idx += 32) {
            const double v = o.dl[idx];
            nv[idx] = v;
            gn = fma(o.g[idx], v, gn);
            nn = fma(v, v, nn);
          }
          gn = warp_sum(gn);
          nn = warp_sum(nn);
          const double step = -gg / gHg;              // dx_u = step * g  (optimizeGradientSearch)
          const double uu = step * step * gg, un = step * gn;
          const double f_error = error;
          double new_f = f_error;
          bool stay = true;
          while (stay) {
            // ComputeDoglegPoint: dx_d = cu * g + cn * dx_n
            const double DeltaSq = Delta * Delta;
            double cu, cn;
            if (DeltaSq < uu) { cu = sqrt(DeltaSq / uu) * step; cn = 0.0; }
            else if (DeltaSq < nn) {                  // ComputeBlend
              const double qa = uu - 2. * un + nn, qb = 2. * (un - uu), qc = uu - DeltaSq;
              const double sq = sqrt(qb * qb - 4 * qa * qc);
              const double tau1 = (-qb + sq) / (2. * qa), tau2 = (-qb - sq) / (2. * qa);
              const double tau = (0.0 <= tau1 && tau1 <= 1.0) ? tau1 : tau2;
              cu = (1. - tau) * step; cn = tau;
            } else { cu = 0.0; cn = 1.0; }
            __syncwarp();
            for (int idx = lane; idx < N * b; idx += 32) o.dl[idx] = fma(cu, o.g[idx], cn * nv[idx]);
            __syncwarp();
            new_f = o.template eval_error<true>();
            n_err++;
            // decrease of the linear model M(0) - M(dx) = -(g.dx + 0.5 dx^T H dx), using H dx_n = -g
            const double dM = -(cu * gg + cn * gn) - 0.5 * (cu * cu * gHg - 2. * cu * cn * gg - cn * cn * gn);
            const double rho = (fabs(f_error - new_f) < 1e-15 || fabs(dM) < 1e-15) ? 0.5 : (f_error - new_f) / dM;
            if (rho >= 0.75) {
              const double dnorm = sqrt(cu * cu * gg + 2. * cu * cn * gn + cn * cn * nn);
              Delta = fmax(Delta, 3.0 * dnorm);
              stay = false;
            } else if (rho >= 0.25) {
              stay = false;
            } else if (rho >= 0.0) {
              if (Delta > 1e-5) Delta = 0.5 * Delta;
              stay = false;
            } else {
              if (Delta > 1e-5) { Delta *= 0.5; stay = true; }
              else {
                for (int idx = lane; idx < N * b; idx += 32) o.dl[idx] = 0.0;
                __syncwarp();
                new_f = f_error;
                stay = false;
              }
            }
          }
          o.accept_step();
          error = new_f;
          iterations++;
        } else if constexpr (!is_lm) {
          // GaussNewtonOptimizer::iterate: solve, retract, error
          n_solve++;
          const bool solved = o.solve(0.0);
          if (!solved) { status |= 16; break; }
          error = o.template eval_error<true>();   // error at retract(values, delta) ...
          o.accept_step();                         // ... which become the values
          n_err++;
          iterations++;
        } else {
          PT_BEGIN();
          o.backup_H(hb);
          PT_END(t_bk);
          bool first = true;
          for (;;) {   // while (!tryLambda(...))
            PT_BEGIN();
            if (!first) o.restore_H(hb);
            PT_END(t_bk);
            first = false;
            bool step_is_successful = false, stopSearchingLambda = false;
            double newError = 0.0;
            n_solve++;
            PT_BEGIN();
            const bool solved = o.solve(lambda);
            PT_END(t_solve);
            if (solved) {
              // linearized cost change = error - linear.error(delta) = -(g.delta) - 0.5 delta^T H delta
              //                        = -0.5 g.delta + 0.5 lambda |delta|^2   (using (H + lambda I) delta = -g)
              double gd = 0.0, dd = 0.0;
              for (int idx = lane; idx < N * b; idx += 32) {
                gd = fma(o.g[idx], o.dl[idx], gd);
                dd = fma(o.dl[idx], o.dl[idx], dd);
              }
              gd = warp_sum(gd);
              dd = warp_sum(dd);
              const double linearizedCostChange = -0.5 * gd + 0.5 * lambda * dd;
              if (linearizedCostChange >= 0.0) {
                PT_BEGIN();
                newError = o.template eval_error<true>();
                PT_END(t_err);
                n_err++;
                const double costChange = error - newError;
                if (linearizedCostChange > 2.220446049250313e-16 * fabs(error)) {
                  const double modelFidelity = costChange / linearizedCostChange;
                  step_is_successful = modelFidelity > minModelFidelity;
                }
                const double minAbsoluteTolerance = relativeErrorTol * error;
                if (fabs(costChange) < minAbsoluteTolerance) stopSearchingLambda = true;
              }
            } else {
              status |= 16;
            }
            if (step_is_successful) {
              o.accept_step();
              error = newError;
              lambda = fmax(lambdaLowerBound, lambda / lambdaFactor);
              iterations++;
              break;
            } else if (!stopSearchingLambda) {
              lambda *= lambdaFactor;
              if (lambda >= lambdaUpperBound) { status |= 8; break; }
            } else {
              break;
            }
          }
        }
        // checkConvergence(relativeErrorTol, absoluteErrorTol, errorTol, currentError, error)
        converged = false;
        why = 0;
        if (error <= errorTol) { converged = true; why = 32; }
        else {
          const double absoluteDecrease = currentError - error;
          const double relativeDecrease = absoluteDecrease / currentError;
          const bool rel = (relativeErrorTol != 0.0) && (relativeDecrease <= relativeErrorTol);
          const bool ab = absoluteDecrease <= absoluteErrorTol;
          converged = rel || ab;
          why = (rel ? 2 : 0) | (ab ? 1 : 0);
        }
      } while (iterations < st.max_iter && !converged);
      if (iterations >= st.max_iter) status |= 4;
      else status |= why;
      if (error > currentError) {   // BatchTrajOptimizer.cpp:297-307: return last_values
        status |= 64;
        step_back = true;
        error = currentError;
      }
    }
    if (step_back) o.step_back();
    // ---- outputs ----
    double* tout = pr.out_traj + prob * TL;
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      tout[idx] = o.xs[i * b + d];
      tout[N * D + idx] = o.xs[i * b + D + d];
    }
    if (lane == 0) {
      if (pr.out_error) pr.out_error[prob] = error;
      if (pr.out_iters) pr.out_iters[prob] = iterations;
      if (pr.out_status) pr.out_status[prob] = status;
    }
    __syncwarp();
    }   // optimizer kernel
  }
  if (lane == 0 && pr.counters && (n_lin | n_solve | n_err)) {
    atomicAdd(pr.counters + 0, n_lin);
    atomicAdd(pr.counters + 1, n_solve);
    atomicAdd(pr.counters + 2, n_err);
#ifdef GPMP2B_PHASE_TIMING
    atomicAdd(pr.counters + 3, (unsigned long long)t_lin);
    atomicAdd(pr.counters + 4, (unsigned long long)t_solve);
    atomicAdd(pr.counters + 5, (unsigned long long)t_err);
    atomicAdd(pr.counters + 6, (unsigned long long)t_bk);
    atomicAdd(pr.counters + 7, (unsigned long long)o.pt_cfg); atomicAdd(pr.counters + 8, (unsigned long long)o.pt_acc);
    atomicAdd(pr.counters + 10, (unsigned long long)o.pt_init);
#endif
  }
}
