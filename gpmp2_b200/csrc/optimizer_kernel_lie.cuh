// optimizer_kernel_lie.cuh -- Pose2Vector (SE(2) x R^n) states: the Pose2MobileArm planners.
//
// Replaces, on top of the shared solve / LM state machine of optimizer_kernel.cuh:
//   GaussianProcessPriorLie<Pose2Vector>::evaluateError        gpmp2/gp/GaussianProcessPriorLie.h:61-86
//   GaussianProcessInterpolatorLie<Pose2Vector>::interpolatePose  gpmp2/gp/GaussianProcessInterpolatorLie.h:64-100
//   ProductDynamicLieGroup retract / localCoordinates / Logmap / Expmap   gpmp2/geometry/ProductDynamicLieGroup.h:84-222
//   PriorFactor<Pose2Vector> (-Local(x, prior), H = I)  and Values::retract   [GTSAM, SURVEY.md App. B.3/B.4]
//   JointLimitFactorPose2Vector                                gpmp2/kinematics/JointLimitFactorPose2Vector.h:66-91
//   the Pose2 Lie formulas (Expmap, Logmap, their derivatives, AdjointMap)     [GTSAM Pose2, SURVEY.md App. B.4]
// The forward kinematics / sphere Jacobians of Pose2MobileArm are config_eval<KIND = 1> in device_model.cuh.
//
// Structure used: every interpolation Jacobian is blockdiag(G_a (3x3), s_a I_n), a = x_i, v_i, x_{i+1}, v_{i+1};
// a lane linearizes one configuration (M = J^T J at the interpolated state) and stages M with the non-zero column
// entries of its four H_a; the entry-parallel lanes (one per entry (r, c) of a D x D block) assemble the ten blocks
// (H_a^T M H_b)[r][c] of the interval as 3 x 3 bilinear forms.  The GP-prior Hessian A^T Q^-1 A uses the same layout.
#pragma once
#include "optimizer_kernel.cuh"
#include "pose2.cuh"

// 1: the plain Pose2MobileArm runs a compile-time chain (config_eval<GEN = false>), the other Pose2Vector robots the general
// one; 0: the general chain for all (A/B switch)
#ifndef GPMP2B_LIE_PLAIN_CHAIN
#define GPMP2B_LIE_PLAIN_CHAIN 1
#endif


template <int D, int NDIM, bool EXTRA = false>
struct LieOpt : public VecOpt<D, NDIM, EXTRA> {
  typedef VecOpt<D, NDIM, EXTRA> Base;
  using Base::b; using Base::BD; using Base::BB; using Base::T;
  using Base::rb; using Base::sdf; using Base::st; using Base::hconst;
  using Base::lane; using Base::N; using Base::K; using Base::C;
  using Base::xs; using Base::g; using Base::dl; using Base::Hd; using Base::Ho; using Base::stage;
  using Base::start_conf; using Base::start_vel; using Base::end_conf; using Base::end_vel;
  using Base::fix_conf; using Base::fix_vel;
  static constexpr bool LIE = true;
  static constexpr int GEOM = 24;   // doubles per interval of the geometry table: P1 (9) | P2 (9) | r (3) | pad
  double* cand;
  const unsigned long long* mask_in = nullptr;   // linearize<true>: sphere masks of this trajectory's configurations (pk_mask) or null = all
  using Base::mask_out;
  double* geom = nullptr;           // per-interval Logmap geometry shared by the GP prior and the interpolator (linearize)

  // layout 0: the full per-trajectory layout (smem_layout, lie); phase-kernel pipeline (pk_kernels.cuh): 2 = xs | dl |
  // cand (error kernel), 4 = xs | g | staging | Ho | Hd (linearize kernel, pk_lie_lin_smem)
  __device__ LieOpt(const KRobot& rb_, const KSdf& sdf_, const KSetting& st_, const double* hc, double* smem, bool = true,
                    int layout = 0)
      : Base(rb_, sdf_, st_, hc, smem, true, layout == 4 ? 2 : layout) {
    if (layout == 0) {
      const SmemLayout L = smem_layout(D, N, true);
      cand = smem + L.cand;
      geom = smem + L.geom;
    } else if (layout == 2) {
      cand = smem + 2 * pk_even(N * b);
    } else {
      int off = 0;
      xs = smem + off; off += pk_even(N * b);
      g = smem + off; off += pk_even(N * b);
      stage = smem + off; off += even_stage_doubles();
      geom = smem + off; off += (N - 1) * GEOM;
      Ho = smem + off; off += (N - 1) * BB;
      Hd = smem + off;
      dl = cand = nullptr;
    }
  }

  __device__ __forceinline__ int even_stage_doubles() const { return (4 * lie_stage_per_config(D) + 32 + 1) & ~1; }

  // ---- geometry of interval (i, i+1): r = Logmap(x_i^-1 x_{i+1}) (pose part), P2 = dr/dx_{i+1} = LogmapDerivative,
  //      P1 = dr/dx_i = -LogmapDerivative * Ad((x_i^-1 x_{i+1})^-1)   (GaussianProcessPriorLie.h:71-80) ----
  template <bool JAC>
  __device__ __forceinline__ void interval_geom(const double* S, int i, double (&r)[3], double (&P1)[9], double (&P2)[9]) const {
    const p2::Pose a{S[i * b], S[i * b + 1], S[i * b + 2]}, c{S[(i + 1) * b], S[(i + 1) * b + 1], S[(i + 1) * b + 2]};
    const p2::Pose btw = p2::between(a, c);
    p2::logmap(btw, r);
    if (JAC) {
      p2::logmap_derivative(r, P2);
      double Ad[9], Tm[9];
      p2::adjoint_of_inverse(btw, Ad);
      p2::mul33(P2, Ad, Tm);
#pragma unroll
      for (int k = 0; k < 9; k++) P1[k] = -Tm[k];
    }
  }

  // configuration (i, j) of state array S: interpolated Pose2Vector, and (JAC) the pose parts G_a of the four
  // interpolation Jacobians (GaussianProcessInterpolatorLie.h:74-95)
  struct QFunL {
    const double* S;
    int o0, o1;
    double px, py, th, w0, w1, w2, w3;
    __device__ __forceinline__ double operator()(int d) const {
      if (d == 0) return px;
      if (d == 1) return py;
      if (d == 2) return th;
      double v = w0 * S[o0 + d];
      v = fma(w1, S[o0 + D + d], v);
      v = fma(w2, S[o1 + d], v);
      v = fma(w3, S[o1 + D + d], v);
      return v;
    }
  };
  template <bool JAC>
  __device__ __forceinline__ QFunL config_state_lie(const double* S, int i, int j, double (&G)[4][9]) const {
    QFunL f;
    f.S = S; f.o0 = i * b; f.o1 = min(i + 1, N - 1) * b;
    const p2::Pose a{S[i * b], S[i * b + 1], S[i * b + 2]};
    if (j == 0) {
      f.px = a.x; f.py = a.y; f.th = a.th;
      f.w0 = 1.0; f.w1 = 0.0; f.w2 = 0.0; f.w3 = 0.0;
      if (JAC) {
#pragma unroll
        for (int k = 0; k < 9; k++) { G[0][k] = (k % 4 == 0) ? 1.0 : 0.0; G[1][k] = 0.0; G[2][k] = 0.0; G[3][k] = 0.0; }
      }
      return f;
    }
    f.w0 = st.gpw[j - 1][0]; f.w1 = st.gpw[j - 1][1]; f.w2 = st.gpw[j - 1][2]; f.w3 = st.gpw[j - 1][3];
    double r[3], P1[9], P2[9];
    interval_geom<JAC>(S, i, r, P1, P2);
    double xi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) xi[k] = fma(f.w3, S[(i + 1) * b + D + k], fma(f.w2, r[k], f.w1 * S[i * b + D + k]));
    const p2::Pose e = p2::expmap(xi);
    const p2::Pose q = p2::compose(a, e);
    f.px = q.x; f.py = q.y; f.th = q.th;
    if (JAC) {
      double E[9], Ade[9], EP1[9], EP2[9];
      p2::expmap_derivative(xi, E);
      p2::adjoint_of_inverse(e, Ade);
      p2::mul33(E, P1, EP1);
      p2::mul33(E, P2, EP2);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        G[0][k] = fma(f.w2, EP1[k], Ade[k]);
        G[1][k] = f.w1 * E[k];
        G[2][k] = f.w2 * EP2[k];
        G[3][k] = f.w3 * E[k];
      }
    }
    return f;
  }

  // Values::retract: Pose2 chart retract x * Pose2(dx, dy, dtheta) + vector add (ProductDynamicLieGroup.h:84-90)
  __device__ void compute_cand() {
    for (int i = lane; i < N; i += 32) {
      const p2::Pose q = p2::compose(p2::Pose{xs[i * b], xs[i * b + 1], xs[i * b + 2]},
                                     p2::Pose{dl[i * b], dl[i * b + 1], dl[i * b + 2]});
      cand[i * b] = q.x; cand[i * b + 1] = q.y; cand[i * b + 2] = q.th;
    }
    for (int idx = lane; idx < N * b; idx += 32)
      if ((idx % b) >= 3) cand[idx] = xs[idx] + dl[idx];
    __syncwarp();
  }
  __device__ void accept_step() {
    compute_cand();
    double* t = xs; xs = cand; cand = t;   // cand keeps the previous values for step_back()
    __syncwarp();
  }
  __device__ void step_back() {
    double* t = xs; xs = cand; cand = t;
    __syncwarp();
  }

  // -Local(x, prior) of the Pose2Vector chart: -(between(x, prior)) for the pose, x - prior for the rest
  __device__ __forceinline__ double prior_err(const double* S, int i, const double* prior, int d) const {
    if (d >= 3) return S[i * b + d] - prior[d];
    const p2::Pose btw = p2::between(p2::Pose{S[i * b], S[i * b + 1], S[i * b + 2]}, p2::Pose{prior[0], prior[1], prior[2]});
    return d == 0 ? -btw.x : (d == 1 ? -btw.y : -btw.th);
  }
  // ---- linearize --------------------------------------------------------------------------------------------
  // Entry-parallel layout used by both the GP-prior and the obstacle part: a lane owns entry (r, c) of EVERY D x D
  // block of an interval's local Hessian over (x_i, v_i, x_{i+1}, v_{i+1}) (entry slot e: index lane + 32 e < D^2).
  // All Jacobians here are block-sparse: a column of blockdiag(P (3x3), s I) has at most 3 non-zeros, at rows
  // (0, 1, 2) for a pose column and at its own row otherwise, so every block entry is a 3 x 3 bilinear form.
  static constexpr int NSLOT = (D * D + 31) / 32;
  static constexpr int HC = 4 * D * 3;               // staged columns of the four interpolation Jacobians
  static constexpr int CSTG = T + HC + D;            // staging per configuration: M (packed), columns, cv
  struct Entry {
    int r, c;                // entry of a D x D block (r = D: no entry in this slot)
    int dxx, dvx, dvv;       // offsets inside a packed diagonal block: (r, c), (D + r, c), (D + r, D + c)
    int orc;                 // offset of (r, c) inside a row-major b x b coupling block
  };
  __device__ __forceinline__ Entry entry_of(int e) const {
    Entry en;
    const int idx = lane + 32 * e;
    en.r = idx < D * D ? idx / D : D;
    en.c = idx < D * D ? idx - en.r * D : 0;
    const int r = min(en.r, D - 1);
    en.dxx = r * (r + 1) / 2 + en.c;
    en.dvx = (D + r) * (D + r + 1) / 2 + en.c;
    en.dvv = en.dvx + D;
    en.orc = r * b + en.c;
    return en;
  }

  // PriorFactor<Pose2Vector>(x_k, fix_conf) + PriorFactor(v_k, fix_vel) on support state k = st.fix_index (VecOpt::fix_pass)
  template <bool GRAD>
  __device__ __forceinline__ double fix_pass_lie(const double* S) {
    const int i = st.fix_index;
    double e = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double ex = prior_err(S, i, fix_conf, d), ev = S[i * b + D + d] - fix_vel[d];
      if (GRAD) { g[i * b + d] += st.conf_prior_w * ex; g[i * b + D + d] += st.vel_prior_w * ev; }
      e += 0.5 * (st.conf_prior_w * ex * ex + st.vel_prior_w * ev * ev);
    }
    return e;
  }

  // ---- linearize: the work is arranged for one warp running alone on its data --
  //        * Logmap geometry (r, P1, P2) of all intervals in ONE lane-parallel pass into a table in shared memory
  //          (`geom`); the GP-prior Hessian, its gradient and every interpolated configuration of the interval read it
  //          (recomputing it warp-uniformly per interval and per configuration cost 23 evaluations per linearization);
  //        * the interpolation Jacobians G_a of a configuration computed once, in the same lane-parallel pass as its
  //          forward kinematics, and kept in registers until the configuration's interval is staged;
  //        * GP-prior gradient with lanes <-> (interval, dof) in two passes over the even / odd intervals;
  //        * obstacle gradient folded into the entry lanes' bilinear forms (no divergent `lanes < D` section).
  //      Measured on config 4 (phase pipeline): 29.3 k -> 20 k warp instructions per linearization, 16.9 -> 12.2 ms per step.
  //      config_state_geom = config_state_lie with the interval's geometry read from the table. ----
  template <bool JAC>
  __device__ __forceinline__ QFunL config_state_geom(const double* S, int i, int j, double (&G)[4][9]) const {
    QFunL f;
    f.S = S; f.o0 = i * b; f.o1 = min(i + 1, N - 1) * b;
    const p2::Pose a{S[i * b], S[i * b + 1], S[i * b + 2]};
    if (j == 0) {
      f.px = a.x; f.py = a.y; f.th = a.th;
      f.w0 = 1.0; f.w1 = 0.0; f.w2 = 0.0; f.w3 = 0.0;
      if (JAC) {
#pragma unroll
        for (int k = 0; k < 9; k++) { G[0][k] = (k % 4 == 0) ? 1.0 : 0.0; G[1][k] = 0.0; G[2][k] = 0.0; G[3][k] = 0.0; }
      }
      return f;
    }
    f.w0 = st.gpw[j - 1][0]; f.w1 = st.gpw[j - 1][1]; f.w2 = st.gpw[j - 1][2]; f.w3 = st.gpw[j - 1][3];
    const double* gm = geom + i * GEOM;
    double xi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) xi[k] = fma(f.w3, S[(i + 1) * b + D + k], fma(f.w2, gm[18 + k], f.w1 * S[i * b + D + k]));
    const p2::Pose e = p2::expmap(xi);
    const p2::Pose q = p2::compose(a, e);
    f.px = q.x; f.py = q.y; f.th = q.th;
    if (JAC) {
      double E[9], Ade[9], P[9], EP[9];
      p2::expmap_derivative(xi, E);
      p2::adjoint_of_inverse(e, Ade);
#pragma unroll
      for (int k = 0; k < 9; k++) P[k] = gm[k];
      p2::mul33(E, P, EP);
#pragma unroll
      for (int k = 0; k < 9; k++) { G[0][k] = fma(f.w2, EP[k], Ade[k]); G[1][k] = f.w1 * E[k]; G[3][k] = f.w3 * E[k]; }
#pragma unroll
      for (int k = 0; k < 9; k++) P[k] = gm[9 + k];
      p2::mul33(E, P, EP);
#pragma unroll
      for (int k = 0; k < 9; k++) G[2][k] = f.w2 * EP[k];
    }
    return f;
  }

  // MASKS (linearize kernel of the phase pipeline): skip the spheres that the error evaluation of these very states found
  // out of reach of their hinge (config_eval<MASKED>); mask_in = null evaluates everything
  template <bool MASKS = false>
  __device__ void linearize() {
    {   // template: end-state prior weights only (the GP-prior Hessian depends on the state here)
      const int n2 = (N * BD + (N - 1) * BB + 1) / 2;
      Base::copy_in(reinterpret_cast<const double2*>(hconst), reinterpret_cast<double2*>(Ho), n2);
      for (int idx = lane; idx < N * b; idx += 32) g[idx] = 0.0;
    }
    // geometry table: lanes <-> intervals
    for (int i = lane; i < N - 1; i += 32) {
      double rr[3], P1r[9], P2r[9];
      interval_geom<true>(xs, i, rr, P1r, P2r);
      double* gm = geom + i * GEOM;
#pragma unroll
      for (int k = 0; k < 9; k++) { gm[k] = P1r[k]; gm[9 + k] = P2r[k]; }
#pragma unroll
      for (int k = 0; k < 3; k++) gm[18 + k] = rr[k];
    }
    __syncwarp();
    // priors and limit hinges: lanes <-> (state, dof)
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      double gx = 0.0, gv = 0.0;
      if (i == 0 || i == N - 1) {
        const double ex = prior_err(xs, i, i == 0 ? start_conf : end_conf, d);
        const double ev = xs[i * b + D + d] - (i == 0 ? start_vel : end_vel)[d];
        gx = (i == 0 ? st.conf_prior_w : st.end_conf_prior_w) * ex;
        gv = st.vel_prior_w * ev;
      }
      if (st.flag_pos_limit && d >= 3) {
        const double p = xs[i * b + d], lo = st.pos_lo[d] + st.pos_th[d], hi = st.pos_hi[d] - st.pos_th[d];
        const double e = p < lo ? lo - p : (p <= hi ? 0.0 : p - hi), h = p < lo ? -1.0 : (p <= hi ? 0.0 : 1.0);
        gx = fma(st.pos_w[d] * h, e, gx);
        Hd[i * BD + d * (d + 1) / 2 + d] += st.pos_w[d] * h * h;
      }
      if (st.flag_vel_limit) {
        const double p = xs[i * b + D + d], lo = -st.vel_lim[d] + st.vel_th[d], hi = st.vel_lim[d] - st.vel_th[d];
        const double e = p < lo ? lo - p : (p <= hi ? 0.0 : p - hi), h = p < lo ? -1.0 : (p <= hi ? 0.0 : 1.0);
        gv = fma(st.vel_w[d] * h, e, gv);
        const int r = D + d;
        Hd[i * BD + r * (r + 1) / 2 + r] += st.vel_w[d] * h * h;
      }
      if (d == 1) gv = fma(st.veh_w, xs[i * b + D + 1], gv);   // VehicleDynamicsFactorPose2Vector (weight 0 = off)
      g[i * b + d] += gx;
      g[i * b + D + d] += gv;
    }
    __syncwarp();
    if constexpr (EXTRA) {   // optional factors of hand-built graphs (optimizer_kernel.cuh: goal_eval, self_eval)
      if (st.goal_enabled == 1) Base::template goal_eval<1, true>([&](int k) { return xs[(N - 1) * b + k]; });
      if (st.goal_enabled == 2) {
        if (lane == 0) Base::template pose_eval<1, true>(N - 1, [&](int k) { return xs[(N - 1) * b + k]; });
        __syncwarp();
      }
      if (st.n_self) {
        for (int i = lane; i < N; i += 32) Base::template self_eval<1, true>(i, [&](int k) { return xs[i * b + k]; });
        __syncwarp();
      }
      if (st.orient_enabled) {
        for (int i = st.orient_first + lane; i <= st.orient_last; i += 32)
          Base::template orient_eval<1, true>(i, [&](int k) { return xs[i * b + k]; });
        __syncwarp();
      }
      if (st.fix_enabled) { fix_pass_lie<true>(xs); __syncwarp(); }
    }

    Entry ent[NSLOT];
#pragma unroll
    for (int e = 0; e < NSLOT; e++) ent[e] = entry_of(e);

    // ---- GP prior factors (GaussianProcessPriorLie.h:61-86): A = [[J1, -dt I, J2, 0], [0, -I, 0, I]] with
    //      J1 = blockdiag(P1, -I), J2 = blockdiag(P2, I), Q^-1 = qi (x) W, W = Qc^-1.  With U_k = J_k^T W:
    //      (x1,x1) q11 U1 J1   (v1,x1) k1 U1^T     (v1,v1) k2 W       k1 = -(dt q11 + q12), k2 = dt^2 q11 + 2 dt q12 + q22
    //      (x1,x2) q11 U1 J2   (x1,v2) q12 U1      (v1,x2) k1 U2^T    (v1,v2) k3 W,   k3 = -(dt q12 + q22)
    //      (x2,x2) q11 U2 J2   (v2,x2) q12 U2^T    (v2,v2) q22 W
    //      A lane owns the same entries of every interval, so consecutive intervals need no barrier. ----
    {
      const double dt = st.delta_t, q11 = st.qi[0][0], q12 = st.qi[0][1], q22 = st.qi[1][1];
      const double k1 = -(dt * q11 + q12), k2 = dt * dt * q11 + 2.0 * dt * q12 + q22, k3 = -(dt * q12 + q22);
#pragma unroll 1
      for (int i = 0; i < N - 1; i++) {
        const double* P1 = geom + i * GEOM;
        const double* P2 = P1 + 9;
        double* Hdi = Hd + i * BD;
        double* Hoi = Ho + i * BB;
        double* Hdn = Hdi + BD;
#pragma unroll
        for (int e = 0; e < NSLOT; e++) {
          const int r = ent[e].r, c = ent[e].c;
          if (r < D) {
            double u1[3], u2[3], u1c, u2c;
            const double* Wr = st.Qc_inv + r * D;
            if (r < 3) {
#pragma unroll
              for (int x = 0; x < 3; x++) {
                u1[x] = fma(P1[6 + r], st.Qc_inv[2 * D + x], fma(P1[3 + r], st.Qc_inv[D + x], P1[r] * st.Qc_inv[x]));
                u2[x] = fma(P2[6 + r], st.Qc_inv[2 * D + x], fma(P2[3 + r], st.Qc_inv[D + x], P2[r] * st.Qc_inv[x]));
              }
              u1c = fma(P1[6 + r], st.Qc_inv[2 * D + c], fma(P1[3 + r], st.Qc_inv[D + c], P1[r] * st.Qc_inv[c]));
              u2c = fma(P2[6 + r], st.Qc_inv[2 * D + c], fma(P2[3 + r], st.Qc_inv[D + c], P2[r] * st.Qc_inv[c]));
            } else {
#pragma unroll
              for (int x = 0; x < 3; x++) { u1[x] = -Wr[x]; u2[x] = Wr[x]; }
              u1c = -Wr[c]; u2c = Wr[c];
            }
            const double w = Wr[c];
            double x11, x12, x22;
            if (c < 3) {
              x11 = fma(u1[2], P1[6 + c], fma(u1[1], P1[3 + c], u1[0] * P1[c]));
              x12 = fma(u1[2], P2[6 + c], fma(u1[1], P2[3 + c], u1[0] * P2[c]));
              x22 = fma(u2[2], P2[6 + c], fma(u2[1], P2[3 + c], u2[0] * P2[c]));
            } else {
              x11 = -u1c; x12 = u1c; x22 = u2c;
            }
            const int tvx = (D + c) * (D + c + 1) / 2 + r, tor = (D + c) * b + r;
            if (r >= c) { Hdi[ent[e].dxx] += q11 * x11; Hdi[ent[e].dvv] += k2 * w; }
            Hdi[tvx] += k1 * u1c;
            Hoi[ent[e].orc] += q11 * x12;
            Hoi[ent[e].orc + D] += q12 * u1c;
            Hoi[tor] += k1 * u2c;
            Hoi[ent[e].orc + D * b + D] += k3 * w;
            if (r >= c) { Hdn[ent[e].dxx] += q11 * x22; Hdn[ent[e].dvv] += q22 * w; }
            Hdn[tvx] += q12 * u2c;
          }
        }
      }
      // gradient u = Q^-1 e, g += A^T u: lanes <-> (interval, dof), the even intervals first, then the odd ones
      // (intervals of one parity share no state)
#pragma unroll 1
      for (int par = 0; par < 2; par++) {
        const int nint = (N - 1 - par + 1) / 2;   // intervals par, par + 2, ...
        for (int idx = lane; idx < nint * D; idx += 32) {
          const int ii = idx / D, r = idx - ii * D, i = 2 * ii + par;
          const double* P1 = geom + i * GEOM;
          const double* P2 = P1 + 9;
          const double* rr = P1 + 18;
          double ux[3], uv_own = 0.0, ux_own = 0.0;
#pragma unroll
          for (int x = 0; x < 3; x++) ux[x] = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) {
            const double rk = (k < 3) ? rr[k] : xs[(i + 1) * b + k] - xs[i * b + k];
            const double ex = rk - dt * xs[i * b + D + k], ev = xs[(i + 1) * b + D + k] - xs[i * b + D + k];
            const double sx = fma(q11, ex, q12 * ev), sv = fma(q12, ex, q22 * ev);
#pragma unroll
            for (int x = 0; x < 3; x++) ux[x] = fma(st.Qc_inv[x * D + k], sx, ux[x]);
            ux_own = fma(st.Qc_inv[r * D + k], sx, ux_own);
            uv_own = fma(st.Qc_inv[r * D + k], sv, uv_own);
          }
          double j1u, j2u;
          if (r < 3) {
            j1u = fma(P1[6 + r], ux[2], fma(P1[3 + r], ux[1], P1[r] * ux[0]));
            j2u = fma(P2[6 + r], ux[2], fma(P2[3 + r], ux[1], P2[r] * ux[0]));
          } else { j1u = -ux_own; j2u = ux_own; }
          g[i * b + r] += j1u;
          g[i * b + D + r] += -dt * ux_own - uv_own;
          g[(i + 1) * b + r] += j2u;
          g[(i + 1) * b + D + r] += uv_own;
        }
        __syncwarp();
      }
    }

    // ---- obstacle factors.  Interval-aligned passes: a pass evaluates 32 / (K + 1) whole intervals (lane ->
    //      (interval slot, j)); the unary factor of the last state rides in a spare lane of pass 0 when there is one.
    //      Every lane of a pass evaluates its configuration AND its interpolation Jacobians once; then, interval by
    //      interval, the K + 1 producer lanes stage (M, the non-zero columns of their four interpolation Jacobians
    //      H_a = blockdiag(G_a, s_a I), cv) and the entry lanes accumulate the ten blocks
    //      (H_a^T M H_b)[r][c] = sum_{x,y} alpha_a[x] M[rho_x][kappa_y] beta_b[y] in registers, flushing once. ----
    const int CI = K + 1;
    const int cap = (even_stage_doubles() - 32) / CSTG;
    const int IPP = max(1, 32 / CI);
    const bool spare = IPP * CI < 32 && CI <= 32;
    int moff[NSLOT][9];
#pragma unroll
    for (int e = 0; e < NSLOT; e++) {
      const int r = min(ent[e].r, D - 1), c = ent[e].c;
#pragma unroll
      for (int x = 0; x < 3; x++)
#pragma unroll
        for (int y = 0; y < 3; y++) {
          const int p = r < 3 ? x : r, q = c < 3 ? y : c;
          const int hi = p > q ? p : q, lo = p > q ? q : p;
          moff[e][x * 3 + y] = hi * (hi + 1) / 2 + lo;
        }
    }
    const int n_int = N - 1;
    const int n_pass = CI <= 32 ? (n_int + IPP - 1) / IPP + ((spare || n_int == 0) ? 0 : 1) : 0;
#pragma unroll 1
    for (int pass = 0; pass < n_pass; pass++) {
      const bool extra = pass * IPP >= n_int;
      const int i0 = pass * IPP;
      const int li = lane / CI, lj = lane - li * CI;
      int ci = i0 + li, cj = lj;
      bool valid = !extra && li < IPP && ci < n_int;
      const bool last_state = extra ? lane == 0 : (spare && pass == 0 && lane == IPP * CI);
      if (last_state) { ci = N - 1; cj = 0; valid = true; }
      double M[T], cv[D], G[4][9], sw[4];
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        sw[a] = 0.0;
#pragma unroll
        for (int k = 0; k < 9; k++) G[a][k] = 0.0;
      }
      unsigned long long sm = ~0ull, wm = ~0ull;
      if (MASKS) {
        if (mask_in) sm = valid ? mask_in[ci * (K + 1) + cj] : 0ull;
        const unsigned lo = __reduce_or_sync(FULL_MASK, (unsigned)sm), hi = __reduce_or_sync(FULL_MASK, (unsigned)(sm >> 32));
        wm = ((unsigned long long)hi << 32) | lo;
      }
      if (valid && wm != 0ull) {
        double e2 = 0.0, es = 0.0;
        const QFunL qf = config_state_geom<true>(xs, ci, cj, G);
        sw[0] = qf.w0; sw[1] = qf.w1; sw[2] = qf.w2; sw[3] = qf.w3;
        // (plain Pose2MobileArm: the compile-time chain; the other Pose2Vector robots: the general one -- kernel-uniform branch)
        if (GPMP2B_LIE_PLAIN_CHAIN && rb.kind == 1)
          config_eval<D, NDIM, 1, true, false, MASKS, false>(rb, sdf, qf, st.epsilon, st.inv_cost_sigma, M, cv, e2, es, nullptr, nullptr, sm, wm);
        else
          config_eval<D, NDIM, 1, true, false, MASKS, true>(rb, sdf, qf, st.epsilon, st.inv_cost_sigma, M, cv, e2, es, nullptr, nullptr, sm, wm);
      }
      const int ns = extra ? 0 : min(IPP, n_int - i0);
#pragma unroll 1
      for (int sl = 0; sl <= ns; sl++) {
        const bool tail = sl == ns;
        if (tail && !(extra || (spare && pass == 0))) break;
        const int i = tail ? N - 1 : i0 + sl;
        const int ncfg = tail ? 1 : CI;
        double acc[NSLOT][10], gacc[NSLOT][4];
#pragma unroll
        for (int e = 0; e < NSLOT; e++) {
#pragma unroll
          for (int t = 0; t < 10; t++) acc[e][t] = 0.0;
#pragma unroll
          for (int t = 0; t < 4; t++) gacc[e][t] = 0.0;
        }
#pragma unroll 1
        for (int j0 = 0; j0 < ncfg; j0 += cap) {
          const int nr = min(cap, ncfg - j0);
          const bool mine = tail ? last_state : (valid && !last_state && li == sl && lj >= j0 && lj < j0 + nr);
          if (mine) {
            double* sp = stage + (tail ? 0 : lj - j0) * CSTG;
#pragma unroll
            for (int m = 0; m < T; m++) sp[m] = M[m];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
              for (int r = 0; r < D; r++)
#pragma unroll
                for (int x = 0; x < 3; x++)
                  sp[T + (a * D + r) * 3 + x] = r < 3 ? G[a][x * 3 + r] : (x == 0 ? sw[a] : 0.0);
#pragma unroll
            for (int d = 0; d < D; d++) sp[T + HC + d] = cv[d];
          }
          __syncwarp();
#pragma unroll 1
          for (int u = 0; u < nr; u++) {
            const double* sp = stage + u * CSTG;
            const double* hc = sp + T;
            const double* cvs = sp + T + HC;
#pragma unroll
            for (int e = 0; e < NSLOT; e++) {
              const int r = ent[e].r, c = ent[e].c;
              if (r < D) {
                double m9[9];
#pragma unroll
                for (int k = 0; k < 9; k++) m9[k] = sp[moff[e][k]];
                double tb[4][3];
#pragma unroll
                for (int bq = 0; bq < 4; bq++) {
                  const double* be = hc + (bq * D + c) * 3;
                  const double b0 = be[0], b1 = be[1], b2 = be[2];
#pragma unroll
                  for (int x = 0; x < 3; x++) tb[bq][x] = fma(m9[x * 3 + 2], b2, fma(m9[x * 3 + 1], b1, m9[x * 3] * b0));
                }
                double al[4][3];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                  const double* ae = hc + (a * D + r) * 3;
                  al[a][0] = ae[0]; al[a][1] = ae[1]; al[a][2] = ae[2];
                }
                auto blk = [&](int a, int bq) { return fma(al[a][2], tb[bq][2], fma(al[a][1], tb[bq][1], al[a][0] * tb[bq][0])); };
                acc[e][0] += blk(0, 0); acc[e][1] += blk(1, 0); acc[e][2] += blk(1, 1);
                acc[e][3] += blk(0, 2); acc[e][4] += blk(0, 3); acc[e][5] += blk(1, 2); acc[e][6] += blk(1, 3);
                acc[e][7] += blk(2, 2); acc[e][8] += blk(3, 2); acc[e][9] += blk(3, 3);
                // gradient rows g_a[r] = sum_x alpha_a[x] cv[rho_x] (flushed by the lanes with c = 0)
                const double c0 = cvs[r < 3 ? 0 : r], c1 = cvs[1], c2 = cvs[2];
#pragma unroll
                for (int a = 0; a < 4; a++) gacc[e][a] += fma(al[a][2], c2, fma(al[a][1], c1, al[a][0] * c0));
              }
            }
          }
          __syncwarp();
        }
        const bool has_next = i < N - 1;
#pragma unroll
        for (int e = 0; e < NSLOT; e++) {
          const int r = ent[e].r, c = ent[e].c;
          if (r < D) {
            double* Hdi = Hd + i * BD;
            if (r >= c) { Hdi[ent[e].dxx] += acc[e][0]; Hdi[ent[e].dvv] += acc[e][2]; }
            Hdi[ent[e].dvx] += acc[e][1];
            if (has_next) {
              double* Hoi = Ho + i * BB;
              double* Hdn = Hdi + BD;
              Hoi[ent[e].orc] += acc[e][3];
              Hoi[ent[e].orc + D] += acc[e][4];
              Hoi[ent[e].orc + D * b] += acc[e][5];
              Hoi[ent[e].orc + D * b + D] += acc[e][6];
              if (r >= c) { Hdn[ent[e].dxx] += acc[e][7]; Hdn[ent[e].dvv] += acc[e][9]; }
              Hdn[ent[e].dvx] += acc[e][8];
            }
            if (c == 0) {
              g[i * b + r] += gacc[e][0];
              g[i * b + D + r] += gacc[e][1];
              if (has_next) { g[(i + 1) * b + r] += gacc[e][2]; g[(i + 1) * b + D + r] += gacc[e][3]; }
            }
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }

  // ---- NonlinearFactorGraph::error at xs (CAND = false) or retract(xs, dl) (CAND = true) ----
  template <bool CAND, bool MASK = false>
  __device__ double eval_error() {
    if (CAND) compute_cand();
    const double* S = CAND ? cand : xs;
    double eacc = 0.0;
    // GP prior: lanes <-> intervals
    for (int i = lane; i < N - 1; i += 32) {
      double r[3], P1[9], P2[9];
      interval_geom<false>(S, i, r, P1, P2);
      double e[2 * D];
#pragma unroll
      for (int k = 0; k < D; k++) {
        const double rk = (k < 3) ? r[k] : S[(i + 1) * b + k] - S[i * b + k];
        e[k] = rk - st.delta_t * S[i * b + D + k];
        e[D + k] = S[(i + 1) * b + D + k] - S[i * b + D + k];
      }
      // e^T (qi (x) W) e = q11 ex.W ex + 2 q12 ex.W ev + q22 ev.W ev,  W = Qc^-1
      double xx = 0.0, xv = 0.0, vv = 0.0;
#pragma unroll
      for (int p = 0; p < D; p++) {
        double wx = 0.0, wv = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) { wx = fma(st.Qc_inv[p * D + q], e[q], wx); wv = fma(st.Qc_inv[p * D + q], e[D + q], wv); }
        xx = fma(e[p], wx, xx); xv = fma(e[p], wv, xv); vv = fma(e[D + p], wv, vv);
      }
      eacc += 0.5 * (st.qi[0][0] * xx + 2.0 * st.qi[0][1] * xv + st.qi[1][1] * vv);
    }
    // priors + limits: lanes <-> (state, dof)
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      if (i == 0 || i == N - 1) {
        const double ex = prior_err(S, i, i == 0 ? start_conf : end_conf, d);
        const double ev = S[i * b + D + d] - (i == 0 ? start_vel : end_vel)[d];
        eacc += 0.5 * ((i == 0 ? st.conf_prior_w : st.end_conf_prior_w) * ex * ex + st.vel_prior_w * ev * ev);
      }
      if (d == 1) { const double vy = S[i * b + D + 1]; eacc = fma(0.5 * st.veh_w * vy, vy, eacc); }   // vehicle dynamics (0 = off)
      if (st.flag_pos_limit && d >= 3) {
        const double p = S[i * b + d], lo = st.pos_lo[d] + st.pos_th[d], hi = st.pos_hi[d] - st.pos_th[d];
        const double e = p < lo ? lo - p : (p <= hi ? 0.0 : p - hi);
        eacc += 0.5 * st.pos_w[d] * e * e;
      }
      if (st.flag_vel_limit) {
        const double p = S[i * b + D + d], lo = -st.vel_lim[d] + st.vel_th[d], hi = st.vel_lim[d] - st.vel_th[d];
        const double e = p < lo ? lo - p : (p <= hi ? 0.0 : p - hi);
        eacc += 0.5 * st.vel_w[d] * e * e;
      }
    }
    if constexpr (EXTRA) {
      if (st.goal_enabled == 1) eacc += Base::template goal_eval<1, false>([&](int k) { return S[(N - 1) * b + k]; });
      if (st.goal_enabled == 2 && lane == 0) eacc += Base::template pose_eval<1, false>(N - 1, [&](int k) { return S[(N - 1) * b + k]; });
      if (st.n_self)
        for (int i = lane; i < N; i += 32) eacc += Base::template self_eval<1, false>(i, [&](int k) { return S[i * b + k]; });
      if (st.orient_enabled)
        for (int i = st.orient_first + lane; i <= st.orient_last; i += 32)
          eacc += Base::template orient_eval<1, false>(i, [&](int k) { return S[i * b + k]; });
      if (st.fix_enabled) eacc += fix_pass_lie<false>(S);
    }
    double e2 = 0.0;
    int chunk;
    double* scratch = Base::err_scratch(chunk);
    __syncwarp();
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      if (cidx < C) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double G[4][9], es = 0.0;
        unsigned long long am = 0ull;
        if (GPMP2B_LIE_PLAIN_CHAIN && rb.kind == 1)
          config_error<D, NDIM, 1, false, MASK, false>(rb, sdf, config_state_lie<false>(S, i, j, G), st.epsilon, st.inv_cost_sigma, e2, es,
                                                       nullptr, nullptr, scratch, chunk, &am);
        else
          config_error<D, NDIM, 1, false, MASK, true>(rb, sdf, config_state_lie<false>(S, i, j, G), st.epsilon, st.inv_cost_sigma, e2, es,
                                                      nullptr, nullptr, scratch, chunk, &am);
        if (MASK) mask_out[cidx] = am;
      }
    }
    return warp_sum(eacc + 0.5 * e2);
  }

  __device__ double collision_cost() {
    double es = 0.0;
    int chunk;
    double* scratch = Base::err_scratch(chunk);
    for (int i = lane; i < N; i += 32) {
      double G[4][9], e2 = 0.0;
      config_error<D, NDIM, 1, false, false, true>(rb, sdf, config_state_lie<false>(xs, i, 0, G), 0.0, 1.0, e2, es, nullptr, nullptr, scratch, chunk);
    }
    return warp_sum(es);
  }

  __device__ void debug_obs(int cidx, double* de, double* dc) {
    const int i = cidx / (K + 1), j = cidx - i * (K + 1);
    double G[4][9], e2 = 0.0, es = 0.0;
    config_error<D, NDIM, 1, true, false, true>(rb, sdf, config_state_lie<false>(xs, i, j, G), st.epsilon, st.inv_cost_sigma, e2, es, de, dc);
  }
};
