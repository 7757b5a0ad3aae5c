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
// a lane linearizes one configuration (M = J^T J at the interpolated state), forms Y_a = M H_a, and the
// entry-parallel lanes assemble (H_a^T M H_b)[r][c] = sum_p H_a[p][r] Y_b[p][c] (3 terms for pose rows, 1 otherwise).
#pragma once
#include "optimizer_kernel.cuh"
#include "pose2.cuh"


template <int D, int NDIM>
struct LieOpt : public VecOpt<D, NDIM> {
  typedef VecOpt<D, NDIM> Base;
  using Base::b; using Base::BD; using Base::BB; using Base::T;
  using Base::rb; using Base::sdf; using Base::st; using Base::hconst;
  using Base::lane; using Base::N; using Base::K; using Base::C;
  using Base::xs; using Base::g; using Base::dl; using Base::Hd; using Base::Ho; using Base::stage;
  using Base::start_conf; using Base::start_vel; using Base::end_conf; using Base::end_vel;
  static constexpr bool LIE = true;
  static constexpr int LSTG = 4 * D * D + 36 + 4 + D;   // per-configuration staging: Y[4][D][D], G[4][9], s[4], cv[D]
  static constexpr int NENT = (2 * BD + BB + 31) / 32;  // output entries per lane of one interval's local Hessian
  double* cand;
  double* tmp;   // 32 doubles after the staging area: interval geometry (r[3], P1[9], P2[9])

  __device__ LieOpt(const KRobot& rb_, const KSdf& sdf_, const KSetting& st_, const double* hc, double* smem)
      : Base(rb_, sdf_, st_, hc, smem, true) {
    const SmemLayout L = smem_layout(D, N, true);
    cand = smem + L.cand;
    tmp = stage + 4 * LSTG;
  }

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
  // (Q^-1)[p][q], Q^-1 = qi (x) Qc^-1
  __device__ __forceinline__ double qinv(int p, int q) const { return st.qi[p / D][q / D] * st.Qc_inv[(p % D) * D + (q % D)]; }
  // GP prior error e[k], k < 2D, of interval i (needs r in tmp[0..2])
  __device__ __forceinline__ double gp_err(const double* S, int i, int k) const {
    if (k >= D) return S[(i + 1) * b + k] - S[i * b + k];   // v_{i+1} - v_i  (k - D + D)
    const double rk = (k < 3) ? tmp[k] : S[(i + 1) * b + k] - S[i * b + k];
    return rk - st.delta_t * S[i * b + D + k];
  }
  // column (a, r) of the GP-prior Jacobian A = [[Jr1, -dt I, Jr2, 0], [0, -I, 0, I]] (2D rows): up to 3 nonzeros
  __device__ __forceinline__ int gp_col(int a, int r, int (&rows)[3], double (&vals)[3]) const {
    if (a == 0 || a == 2) {
      if (r < 3) {
        const double* P = tmp + (a == 0 ? 3 : 12);
#pragma unroll
        for (int p = 0; p < 3; p++) { rows[p] = p; vals[p] = P[p * 3 + r]; }
        return 3;
      }
      rows[0] = r; vals[0] = (a == 0) ? -1.0 : 1.0;
      return 1;
    }
    if (a == 1) { rows[0] = r; vals[0] = -st.delta_t; rows[1] = D + r; vals[1] = -1.0; return 2; }
    rows[0] = D + r; vals[0] = 1.0;
    return 1;
  }

  // descriptor of output entry en (< 2 BD + BB) of one interval's local Hessian:
  // bits 0-1 a, 2-4 r, 5-6 b, 7-9 c, 10-11 kind (0 Hd[i], 1 Ho[i], 2 Hd[i+1]), 12.. offset inside the target block
  __device__ __forceinline__ int entry_desc(int en) const {
    int kind, off, ra, ca;
    if (en < BD || en >= BD + BB) {
      kind = en < BD ? 0 : 2;
      off = en < BD ? en : en - BD - BB;
      ra = (int)((sqrtf(8.0f * (float)off + 1.0f) - 1.0f) * 0.5f);
      if (ra * (ra + 1) / 2 > off) ra--;
      if ((ra + 1) * (ra + 2) / 2 <= off) ra++;
      ca = off - ra * (ra + 1) / 2;
      if (kind == 2) { ra += b; ca += b; }
    } else {
      kind = 1; off = en - BD;
      ra = off / b; ca = b + off % b;
    }
    return (ra / D) | ((ra % D) << 2) | ((ca / D) << 5) | ((ca % D) << 7) | (kind << 10) | (off << 12);
  }

  // ---- linearize ----
  __device__ void linearize() {
    {   // template: end-state prior weights only (the GP-prior Hessian depends on the state here)
      const int n2 = (N * BD + (N - 1) * BB + 1) / 2;
      const double2* src = reinterpret_cast<const double2*>(hconst);
      double2* dst = reinterpret_cast<double2*>(Ho);
      for (int idx = lane; idx < n2; idx += 32) dst[idx] = __ldg(src + idx);
      for (int idx = lane; idx < N * b; idx += 32) g[idx] = 0.0;
    }
    int desc[NENT];
#pragma unroll
    for (int t = 0; t < NENT; t++) desc[t] = entry_desc(min(lane + 32 * t, 2 * BD + BB - 1));
    __syncwarp();
    // priors and limit hinges: lanes <-> (state, dof)
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      double gx = 0.0, gv = 0.0;
      if (i == 0 || i == N - 1) {
        const double ex = prior_err(xs, i, i == 0 ? start_conf : end_conf, d);
        const double ev = xs[i * b + D + d] - (i == 0 ? start_vel : end_vel)[d];
        gx = st.conf_prior_w * ex;
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
      g[i * b + d] += gx;
      g[i * b + D + d] += gv;
    }
    __syncwarp();
    // GP prior factors, one interval at a time: H += A^T Q^-1 A, g += A^T Q^-1 e
#pragma unroll 1
    for (int i = 0; i < N - 1; i++) {
      {
        double r[3], P1[9], P2[9];
        interval_geom<true>(xs, i, r, P1, P2);
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 3; k++) tmp[k] = r[k];
#pragma unroll
          for (int k = 0; k < 9; k++) { tmp[3 + k] = P1[k]; tmp[12 + k] = P2[k]; }
        }
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < NENT; t++) {
        if (lane + 32 * t < 2 * BD + BB) {
          const int ds = desc[t];
          int ra[3], rc[3];
          double va[3], vc[3];
          const int na = gp_col(ds & 3, (ds >> 2) & 7, ra, va), nc = gp_col((ds >> 5) & 3, (ds >> 7) & 7, rc, vc);
          double acc = 0.0;
          for (int x = 0; x < na; x++)
            for (int y = 0; y < nc; y++) acc = fma(va[x] * qinv(ra[x], rc[y]), vc[y], acc);
          const int kind = (ds >> 10) & 3, off = ds >> 12;
          double* tgt = kind == 0 ? Hd + i * BD : (kind == 1 ? Ho + i * BB : Hd + (i + 1) * BD);
          tgt[off] += acc;
        }
      }
      if (lane < 4 * D) {   // gradient entries (a, r)
        const int a = lane / D, r = lane - a * D;
        int ra[3];
        double va[3];
        const int na = gp_col(a, r, ra, va);
        double acc = 0.0;
        for (int x = 0; x < na; x++) {
          double u = 0.0;
          for (int k = 0; k < 2 * D; k++) u = fma(qinv(ra[x], k), gp_err(xs, i, k), u);
          acc = fma(va[x], u, acc);
        }
        g[(i + a / 2) * b + (a & 1) * D + r] += acc;
      }
      __syncwarp();
    }

    // obstacle factors: configuration-parallel, then entry-parallel in rounds of 4 configurations
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      const bool valid = cidx < C;
      double M[T], cv[D], G[4][9], sw[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) cv[d] = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < 9; k++) G[a][k] = 0.0;
      if (valid) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double e2 = 0.0, es = 0.0;
        const QFunL qf = config_state_lie<true>(xs, i, j, G);
        sw[0] = qf.w0; sw[1] = qf.w1; sw[2] = qf.w2; sw[3] = qf.w3;
        config_eval<D, NDIM, 1, true, false>(rb, sdf, qf, st.epsilon, st.inv_cost_sigma, M, cv, e2, es, nullptr, nullptr);
      }
      int ri = c0 / (K + 1), rj = c0 - ri * (K + 1);
#pragma unroll 1
      for (int round = 0; round < 8; round++) {
        const int ci0 = c0 + round * 4;
        if (ci0 >= C) break;
        if ((lane >> 2) == round && valid) {
          double* sp = stage + (lane & 3) * LSTG;
          // Y_a = M H_a, H_a = blockdiag(G_a, s_a I)
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int p = 0; p < D; p++)
#pragma unroll
              for (int c = 0; c < D; c++) {
                double v;
                if (c < 3) {
                  v = 0.0;
#pragma unroll
                  for (int q = 0; q < 3; q++) {
                    const int hi = p > q ? p : q, lo = p > q ? q : p;
                    v = fma(M[hi * (hi + 1) / 2 + lo], G[a][q * 3 + c], v);
                  }
                } else {
                  const int hi = p > c ? p : c, lo = p > c ? c : p;
                  v = sw[a] * M[hi * (hi + 1) / 2 + lo];
                }
                sp[(a * D + p) * D + c] = v;
              }
#pragma unroll
          for (int a = 0; a < 4; a++) {
#pragma unroll
            for (int k = 0; k < 9; k++) sp[4 * D * D + a * 9 + k] = G[a][k];
            sp[4 * D * D + 36 + a] = sw[a];
          }
#pragma unroll
          for (int d = 0; d < D; d++) sp[4 * D * D + 40 + d] = cv[d];
        }
        __syncwarp();
        const int nt = min(4, C - ci0);
        int t = 0;
#pragma unroll 1
        while (t < nt) {
          // segment = configurations of interval ri present in this round
          const int jn = min(K + 1 - rj, nt - t);
          double acc[NENT], gacc = 0.0;
#pragma unroll
          for (int x = 0; x < NENT; x++) acc[x] = 0.0;
          const bool last_state = ri == N - 1;
#pragma unroll 1
          for (int u = 0; u < jn; u++) {
            const double* sp = stage + (t + u) * LSTG;
            const double* Gs = sp + 4 * D * D;
            const double* ss = Gs + 36;
#pragma unroll
            for (int x = 0; x < NENT; x++) {
              if (lane + 32 * x < 2 * BD + BB) {
                const int ds = desc[x], a = ds & 3, r = (ds >> 2) & 7, bb = (ds >> 5) & 3, c = (ds >> 7) & 7;
                const double* Yb = sp + bb * D * D;
                double v;
                if (r < 3) v = fma(Gs[a * 9 + 6 + r], Yb[2 * D + c], fma(Gs[a * 9 + 3 + r], Yb[D + c], Gs[a * 9 + r] * Yb[c]));
                else v = ss[a] * Yb[r * D + c];
                acc[x] += v;
              }
            }
            if (lane < 4 * D) {
              const int a = lane / D, r = lane - a * D;
              const double* cvs = sp + 4 * D * D + 40;
              gacc += (r < 3) ? fma(Gs[a * 9 + 6 + r], cvs[2], fma(Gs[a * 9 + 3 + r], cvs[1], Gs[a * 9 + r] * cvs[0]))
                              : ss[a] * cvs[r];
            }
          }
#pragma unroll
          for (int x = 0; x < NENT; x++) {
            if (lane + 32 * x < 2 * BD + BB) {
              const int kind = (desc[x] >> 10) & 3, off = desc[x] >> 12;
              if (kind == 0) Hd[ri * BD + off] += acc[x];
              else if (!last_state) {
                if (kind == 1) Ho[ri * BB + off] += acc[x];
                else Hd[(ri + 1) * BD + off] += acc[x];
              }
            }
          }
          if (lane < 4 * D) {
            const int a = lane / D, r = lane - a * D;
            if (a < 2 || !last_state) g[(ri + a / 2) * b + (a & 1) * D + r] += gacc;
          }
          t += jn; rj += jn;
          if (rj > K) { rj = 0; ri++; }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }

  // ---- NonlinearFactorGraph::error at xs (CAND = false) or retract(xs, dl) (CAND = true) ----
  template <bool CAND>
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
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < 2 * D; p++) {
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < 2 * D; q++) u = fma(qinv(p, q), e[q], u);
        acc = fma(e[p], u, acc);
      }
      eacc += 0.5 * acc;
    }
    // priors + limits: lanes <-> (state, dof)
    for (int idx = lane; idx < N * D; idx += 32) {
      const int i = idx / D, d = idx - i * D;
      if (i == 0 || i == N - 1) {
        const double ex = prior_err(S, i, i == 0 ? start_conf : end_conf, d);
        const double ev = S[i * b + D + d] - (i == 0 ? start_vel : end_vel)[d];
        eacc += 0.5 * (st.conf_prior_w * ex * ex + st.vel_prior_w * ev * ev);
      }
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
    double e2 = 0.0;
    int chunk;
    double* scratch = Base::err_scratch(chunk);
    __syncwarp();
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int cidx = c0 + lane;
      if (cidx < C) {
        const int i = cidx / (K + 1), j = cidx - i * (K + 1);
        double G[4][9], es = 0.0;
        config_error<D, NDIM, 1, false>(rb, sdf, config_state_lie<false>(S, i, j, G), st.epsilon, st.inv_cost_sigma, e2, es,
                                        nullptr, nullptr, scratch, chunk);
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
      config_error<D, NDIM, 1, false>(rb, sdf, config_state_lie<false>(xs, i, 0, G), 0.0, 1.0, e2, es, nullptr, nullptr, scratch, chunk);
    }
    return warp_sum(es);
  }

  __device__ void debug_obs(int cidx, double* de, double* dc) {
    const int i = cidx / (K + 1), j = cidx - i * (K + 1);
    double G[4][9], e2 = 0.0, es = 0.0;
    config_error<D, NDIM, 1, true>(rb, sdf, config_state_lie<false>(xs, i, j, G), st.epsilon, st.inv_cost_sigma, e2, es, de, dc);
  }
};
