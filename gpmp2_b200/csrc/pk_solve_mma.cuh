// pk_solve_mma.cuh -- solve kernel of the phase-kernel pipeline (pk_kernels.cuh) on the FP64 tensor cores:
// a TWO-WARP block per trajectory assembles H and g from the M-list of the linearize kernel and solves
// (H + lambda I) delta = -g with the DMMA block-tridiagonal Cholesky of mma_solve.cuh (one elimination chain per warp).
//
// Replaces pk_solve_kernel (scalar two-sided solver, ~74 k warp instructions per solve, 38 ms of the 67 ms step on the
// headline workload); same inputs and outputs: pk_state (states in, delta + linearized cost change + solved flag out).
//
// What it restates: the entry-parallel half of NonlinearFactorGraph::linearize for the graph of
// internal::BatchTrajOptimize (gpmp2/planner/BatchTrajOptimizer-inl.h:19-84) --
//   PriorFactor on x_0, v_0, x_T, v_T (BatchTrajOptimizer-inl.h:41-48),
//   GaussianProcessPriorLinear between consecutive states (gpmp2/gp/GaussianProcessPriorLinear.h:57-83),
//   ObstacleSDFFactor / ObstacleSDFFactorGP Hessians (w w^T) (x) M from the per-configuration M = J^T J of the
//   linearize kernel (gpmp2/obstacle/ObstacleSDFFactorGP-inl.h:18-75, weights of
//   gpmp2/gp/GaussianProcessInterpolatorLinear.h:62-96),
//   joint / velocity limit hinges (gpmp2/kinematics/JointLimitFactorVector.h:62-79, VelocityLimitFactorVector.h:62-79)
// -- and GTSAM's damped linear solve inside LevenbergMarquardtOptimizer::tryLambda.
//
// Shared memory per trajectory (doubles): g | dl (holds the states during assembly) | scratch | Ho (N-1 row-major
// b x b blocks H_{i,i+1}) | Hd (N packed-lower b x b blocks) = 27.2 KB for the WAM problem -> 8 blocks = 16 warps per SM.
// H is never read from a template: the constant GP-prior / prior part is a handful of scalars times Qc^-1[p][q],
// added on the fly by the lane that owns entry (p, q), so every H entry is stored exactly once.
#pragma once
#include <math_constants.h>
#include <cstdio>
#include "device_model.cuh"
#include "mma_solve.cuh"

namespace pkm {

// ---- gradient of the priors, GP priors and limit hinges: plain stores into g (VecOpt::state_pass<false, true> with
//      the limit curvature left to limit_curvature()); xs: the states, [i * b + d] = x, [i * b + D + d] = v ----
template <int D>
__device__ __forceinline__ void state_gradient(const KSetting& st, const double* xs, double* g, int N, const double* start_conf,
                                               const double* start_vel, const double* end_conf, const double* end_vel, int tid, int nth) {
  constexpr int b = 2 * D;
  const double dt = st.delta_t;
  const double q11 = st.qi[0][0], q12 = st.qi[0][1], q22 = st.qi[1][1];
#pragma unroll 1
  for (int idx = tid; idx < N * D; idx += nth) {
    const int i = idx / D, d = idx - i * D;
    double gx = 0.0, gv = 0.0;
    if (i < N - 1) {   // interval (i, i+1): e = Phi s_i - s_{i+1}; u = Q^-1 e; g_i += Phi^T u
      double ux, uv;
      if (st.qc_identity) {
        const double exd = (xs[i * b + d] + dt * xs[i * b + D + d]) - xs[(i + 1) * b + d];
        const double evd = xs[i * b + D + d] - xs[(i + 1) * b + D + d];
        ux = fma(q11, exd, q12 * evd);
        uv = fma(q12, exd, q22 * evd);
      } else {
        ux = 0.0; uv = 0.0;
#pragma unroll 1
        for (int k = 0; k < D; k++) {
          const double qc = st.Qc_inv[d * D + k];
          const double ex = (xs[i * b + k] + dt * xs[i * b + D + k]) - xs[(i + 1) * b + k];
          const double ev = xs[i * b + D + k] - xs[(i + 1) * b + D + k];
          ux = fma(qc, fma(q11, ex, q12 * ev), ux);
          uv = fma(qc, fma(q12, ex, q22 * ev), uv);
        }
      }
      gx += ux;
      gv += fma(dt, ux, uv);
    }
    if (i > 0) {       // interval (i-1, i): g_i -= u
      double ux, uv;
      if (st.qc_identity) {
        const double ex = (xs[(i - 1) * b + d] + dt * xs[(i - 1) * b + D + d]) - xs[i * b + d];
        const double ev = xs[(i - 1) * b + D + d] - xs[i * b + D + d];
        ux = fma(q11, ex, q12 * ev);
        uv = fma(q12, ex, q22 * ev);
      } else {
        ux = 0.0; uv = 0.0;
#pragma unroll 1
        for (int k = 0; k < D; k++) {
          const double qc = st.Qc_inv[d * D + k];
          const double ex = (xs[(i - 1) * b + k] + dt * xs[(i - 1) * b + D + k]) - xs[i * b + k];
          const double ev = xs[(i - 1) * b + D + k] - xs[i * b + D + k];
          ux = fma(qc, fma(q11, ex, q12 * ev), ux);
          uv = fma(qc, fma(q12, ex, q22 * ev), uv);
        }
      }
      gx -= ux;
      gv -= uv;
    }
    if (i == 0 || i == N - 1) {   // PriorFactor on x_i, v_i
      const double pc = (i == 0 ? start_conf : end_conf)[d], pv = (i == 0 ? start_vel : end_vel)[d];
      const double cw = i == 0 ? st.conf_prior_w : st.end_conf_prior_w;
      gx = fma(cw, xs[i * b + d] - pc, gx);
      gv = fma(st.vel_prior_w, xs[i * b + D + d] - pv, gv);
    }
    if (st.flag_pos_limit) {
      const double p = xs[i * b + d], lo = st.pos_lo[d] + st.pos_th[d], hi = st.pos_hi[d] - st.pos_th[d];
      if (p < lo) gx = fma(-st.pos_w[d], lo - p, gx);
      else if (p > hi) gx = fma(st.pos_w[d], p - hi, gx);
    }
    if (st.flag_vel_limit) {
      const double p = xs[i * b + D + d], lo = -st.vel_lim[d] + st.vel_th[d], hi = st.vel_lim[d] - st.vel_th[d];
      if (p < lo) gv = fma(-st.vel_w[d], lo - p, gv);
      else if (p > hi) gv = fma(st.vel_w[d], p - hi, gv);
    }
    g[i * b + d] = gx;
    g[i * b + D + d] = gv;
  }
}

// curvature of the active limit hinges: w on the diagonal of Hd
template <int D>
__device__ __forceinline__ void limit_curvature(const KSetting& st, const double* xs, double* Hd, int i_begin, int i_end, int tid, int nth) {
  constexpr int b = 2 * D, BD = b * (b + 1) / 2;
  if (!st.flag_pos_limit && !st.flag_vel_limit) return;
#pragma unroll 1
  for (int idx = i_begin * D + tid; idx < i_end * D; idx += nth) {
    const int i = idx / D, d = idx - i * D;
    if (st.flag_pos_limit) {
      const double p = xs[i * b + d], lo = st.pos_lo[d] + st.pos_th[d], hi = st.pos_hi[d] - st.pos_th[d];
      if (p < lo || p > hi) Hd[i * BD + d * (d + 1) / 2 + d] += st.pos_w[d];
    }
    if (st.flag_vel_limit) {
      const double p = xs[i * b + D + d], lo = -st.vel_lim[d] + st.vel_th[d], hi = st.vel_lim[d] - st.vel_th[d];
      const int r = D + d;
      if (p < lo || p > hi) Hd[i * BD + r * (r + 1) / 2 + r] += st.vel_w[d];
    }
  }
}

// ---- H and the obstacle part of g for the intervals [i_begin, i_end) by one warp.  Lane m < T owns entry (p, q),
//      p >= q, of every symmetric D x D sub-block; lanes d < D own component d of the gradient.  Every H entry of the
//      blocks i_begin .. i_end - 1 and of the coupling blocks is stored once (constant part + obstacle part); the
//      contribution to block i_end (the right state of the last interval) is returned in `carry` for the caller.
//      KS: obs_check_inter at compile time (static GP-weight operands), or -1: read K at run time. ----
struct Carry { double xx, xv, vv, gx, gv; };

// the contribution of interval i to its RIGHT state (block i + 1): sums over the interpolated configurations only
template <int D, int KS>
__device__ __forceinline__ Carry right_state_carry(const KSetting& st, const double* __restrict__ ml, int RS, int K, int i, int lane) {
  constexpr int T = D * (D + 1) / 2;
  const bool hlane = lane < T, glane = lane < D;
  const int CI = K + 1;
  const double* rowM = ml + (hlane ? lane : 0);
  const double* rowG = ml + T + (glane ? lane : 0);
  Carry cy; cy.xx = 0.0; cy.xv = 0.0; cy.vv = 0.0; cy.gx = 0.0; cy.gv = 0.0;
  if (KS > 0) {   // static K: all loads in flight together
    constexpr int KA = KS > 0 ? KS : 1, RSC = (T + D + 1) & ~1;
    const double* bm = rowM + (size_t)i * ((KA + 1) * RSC);
    const double* bg = rowG + (size_t)i * ((KA + 1) * RSC);
    double v[KA], gv[KA];
#pragma unroll
    for (int j = 1; j <= KA; j++) { v[j - 1] = __ldg(bm + j * RSC); gv[j - 1] = __ldg(bg + j * RSC); }
#pragma unroll
    for (int j = 1; j <= KA; j++) {
      cy.xx = fma(st.gpww[j - 1][7], v[j - 1], cy.xx); cy.xv = fma(st.gpww[j - 1][8], v[j - 1], cy.xv); cy.vv = fma(st.gpww[j - 1][9], v[j - 1], cy.vv);
      cy.gx = fma(st.gpw[j - 1][2], gv[j - 1], cy.gx); cy.gv = fma(st.gpw[j - 1][3], gv[j - 1], cy.gv);
    }
    return cy;
  }
#pragma unroll 1
  for (int j = 1; j <= K; j++) {
    const double v = __ldg(rowM + (size_t)(i * CI + j) * RS), gvv = __ldg(rowG + (size_t)(i * CI + j) * RS);
    cy.xx = fma(st.gpww[j - 1][7], v, cy.xx); cy.xv = fma(st.gpww[j - 1][8], v, cy.xv); cy.vv = fma(st.gpww[j - 1][9], v, cy.vv);
    cy.gx = fma(st.gpw[j - 1][2], gvv, cy.gx); cy.gv = fma(st.gpw[j - 1][3], gvv, cy.gv);
  }
  return cy;
}

template <int D, int KS>
__device__ __forceinline__ Carry assemble_intervals(const KSetting& st, const double* __restrict__ ml, int RS, double* Hd, double* Ho,
                                                    double* g, int N, int K, int i_begin, int i_end, Carry cy, int lane) {
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b, T = D * (D + 1) / 2;
  // (p, q) of packed entry m = lane
  int p = (int)((sqrtf(8.0f * (float)lane + 1.0f) - 1.0f) * 0.5f);
  if (p * (p + 1) / 2 > lane) p--;
  if ((p + 1) * (p + 2) / 2 <= lane) p++;
  const int q = lane - p * (p + 1) / 2;
  const bool hlane = lane < T, glane = lane < D, offd = p != q;
  const int pc = hlane ? p : 0, qc = hlane ? q : 0;
  const double qpq = st.Qc_inv[pc * D + qc], qqp = st.Qc_inv[qc * D + pc];
  const int dxx = pc * (pc + 1) / 2 + qc, dvx1 = (D + pc) * (D + pc + 1) / 2 + qc, dvx2 = (D + qc) * (D + qc + 1) / 2 + pc,
            dvv = (D + pc) * (D + pc + 1) / 2 + D + qc;
  const int o1 = pc * b + qc, o2 = qc * b + pc;
  const int CI = K + 1;
  const double* rowM = ml + (hlane ? lane : 0);
  const double* rowG = ml + T + (glane ? lane : 0);
  constexpr int KA = KS > 0 ? KS : 1;
  constexpr int RSC = (T + D + 1) & ~1;                // = pk_row_stride(D): row stride at compile time -> immediate offsets
  constexpr int ISTR = (KA + 1) * RSC;                 // doubles per interval (static-K path)
  double nv[KA + 1], ng[KA + 1];
  if (KS > 0 && i_begin < i_end) {
    const double* bm = rowM + (size_t)i_begin * ISTR;
    const double* bg = rowG + (size_t)i_begin * ISTR;
#pragma unroll
    for (int j = 0; j <= KS; j++) { nv[j] = __ldg(bm + j * RSC); ng[j] = __ldg(bg + j * RSC); }
  }
#pragma unroll 1
  for (int i = i_begin; i < i_end; i++) {
    double a0xx = cy.xx, a0xv = cy.xv, a0vv = cy.vv, a1xx = 0, a1xv = 0, a1vv = 0;
    double oxx = 0, oxv = 0, ovx = 0, ovv = 0;
    double g0x = cy.gx, g0v = cy.gv, g1x = 0, g1v = 0;
    if (KS > 0) {
      double val[KA + 1], gval[KA + 1];
#pragma unroll
      for (int j = 0; j <= KS; j++) { val[j] = nv[j]; gval[j] = ng[j]; }
      const int in = min(i + 1, i_end - 1);            // prefetch the next interval's rows (the last one reloads itself)
      const double* bm = rowM + (size_t)in * ISTR;
      const double* bg = rowG + (size_t)in * ISTR;
#pragma unroll
      for (int j = 0; j <= KS; j++) { nv[j] = __ldg(bm + j * RSC); ng[j] = __ldg(bg + j * RSC); }
      a0xx += val[0]; g0x += gval[0];
#pragma unroll
      for (int j = 1; j <= KS; j++) {
        a0xx = fma(st.gpww[j - 1][0], val[j], a0xx); a0xv = fma(st.gpww[j - 1][1], val[j], a0xv); a0vv = fma(st.gpww[j - 1][2], val[j], a0vv);
        oxx = fma(st.gpww[j - 1][3], val[j], oxx);   oxv = fma(st.gpww[j - 1][4], val[j], oxv);
        ovx = fma(st.gpww[j - 1][5], val[j], ovx);   ovv = fma(st.gpww[j - 1][6], val[j], ovv);
        a1xx = fma(st.gpww[j - 1][7], val[j], a1xx); a1xv = fma(st.gpww[j - 1][8], val[j], a1xv); a1vv = fma(st.gpww[j - 1][9], val[j], a1vv);
        g0x = fma(st.gpw[j - 1][0], gval[j], g0x); g0v = fma(st.gpw[j - 1][1], gval[j], g0v);
        g1x = fma(st.gpw[j - 1][2], gval[j], g1x); g1v = fma(st.gpw[j - 1][3], gval[j], g1v);
      }
    } else {
      a0xx += __ldg(rowM + (size_t)(i * CI) * RS); g0x += __ldg(rowG + (size_t)(i * CI) * RS);
#pragma unroll 1
      for (int j = 1; j <= K; j++) {
        const double v = __ldg(rowM + (size_t)(i * CI + j) * RS), gvv = __ldg(rowG + (size_t)(i * CI + j) * RS);
        const double* ww = st.gpww[j - 1];
        const double* w1 = st.gpw[j - 1];
        a0xx = fma(ww[0], v, a0xx); a0xv = fma(ww[1], v, a0xv); a0vv = fma(ww[2], v, a0vv);
        oxx = fma(ww[3], v, oxx);   oxv = fma(ww[4], v, oxv);   ovx = fma(ww[5], v, ovx);   ovv = fma(ww[6], v, ovv);
        a1xx = fma(ww[7], v, a1xx); a1xv = fma(ww[8], v, a1xv); a1vv = fma(ww[9], v, a1vv);
        g0x = fma(w1[0], gvv, g0x); g0v = fma(w1[1], gvv, g0v); g1x = fma(w1[2], gvv, g1x); g1v = fma(w1[3], gvv, g1v);
      }
    }
    if (hlane) {
      // coupling block H_{i,i+1}: GP prior s12 (x) Qc^-1 + obstacle part, both mirror entries
      double* Hoi = Ho + i * BB;
      DBG_IDX(i, N - 1, "coupling block"); DBG_IDX(o1 + D * b + D, BB, "Ho entry"); DBG_IDX(o2 + D * b + D, BB, "Ho entry");
      Hoi[o1] = fma(st.s12[0][0], qpq, oxx);             Hoi[o1 + D] = fma(st.s12[0][1], qpq, oxv);
      Hoi[o1 + D * b] = fma(st.s12[1][0], qpq, ovx);     Hoi[o1 + D * b + D] = fma(st.s12[1][1], qpq, ovv);
      if (offd) {
        Hoi[o2] = fma(st.s12[0][0], qqp, oxx);           Hoi[o2 + D] = fma(st.s12[0][1], qqp, oxv);
        Hoi[o2 + D * b] = fma(st.s12[1][0], qqp, ovx);   Hoi[o2 + D * b + D] = fma(st.s12[1][1], qqp, ovv);
      }
      // diagonal block i: s11 (it has a successor) + s22 (if it has a predecessor) + the priors of state 0
      {
        const double t00 = st.s11[0][0] + (i > 0 ? st.s22[0][0] : 0.0), t10 = st.s11[1][0] + (i > 0 ? st.s22[1][0] : 0.0),
                     t11 = st.s11[1][1] + (i > 0 ? st.s22[1][1] : 0.0);
        double* Hdi = Hd + i * BD;
        DBG_IDX(i, N, "diagonal block"); DBG_IDX(dvv, BD, "Hd entry"); DBG_IDX(dvx2, BD, "Hd entry");
        double vxx = fma(t00, qpq, a0xx), vvv = fma(t11, qpq, a0vv);
        if (i == 0 && !offd) { vxx += st.conf_prior_w; vvv += st.vel_prior_w; }
        Hdi[dxx] = vxx;
        Hdi[dvx1] = fma(t10, qpq, a0xv);
        if (offd) Hdi[dvx2] = fma(t10, qqp, a0xv);
        Hdi[dvv] = vvv;
      }
    }
    if (glane) {
      double* gi = g + i * b + lane;
      gi[0] += g0x;
      gi[D] += g0v;
    }
    cy.xx = a1xx; cy.xv = a1xv; cy.vv = a1vv; cy.gx = g1x; cy.gv = g1v;
  }
  return cy;
}

// the last diagonal block of a range: constant part + carry (+ the unary factor and the priors of the last state)
template <int D>
__device__ __forceinline__ void store_block(const KSetting& st, const double* __restrict__ ml, int RS, double* Hd, double* g, int N, int K,
                                            int i, const Carry& cy, int lane) {
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, T = D * (D + 1) / 2;
  int p = (int)((sqrtf(8.0f * (float)lane + 1.0f) - 1.0f) * 0.5f);
  if (p * (p + 1) / 2 > lane) p--;
  if ((p + 1) * (p + 2) / 2 <= lane) p++;
  const int q = lane - p * (p + 1) / 2;
  const bool hlane = lane < T, glane = lane < D, offd = p != q;
  const int pc = hlane ? p : 0, qc = hlane ? q : 0;
  const double qpq = st.Qc_inv[pc * D + qc], qqp = st.Qc_inv[qc * D + pc];
  const int dxx = pc * (pc + 1) / 2 + qc, dvx1 = (D + pc) * (D + pc + 1) / 2 + qc, dvx2 = (D + qc) * (D + qc + 1) / 2 + pc,
            dvv = (D + pc) * (D + pc + 1) / 2 + D + qc;
  const bool last = i == N - 1;
  double uxx = 0.0, ugx = 0.0;
  if (last) {   // unary obstacle factor of the last support state
    const int CI = K + 1;
    uxx = hlane ? __ldg(ml + (size_t)((N - 1) * CI) * RS + lane) : 0.0;
    ugx = glane ? __ldg(ml + (size_t)((N - 1) * CI) * RS + T + lane) : 0.0;
  }
  if (hlane) {
    const double t00 = (last ? 0.0 : st.s11[0][0]) + (i > 0 ? st.s22[0][0] : 0.0), t10 = (last ? 0.0 : st.s11[1][0]) + (i > 0 ? st.s22[1][0] : 0.0),
                 t11 = (last ? 0.0 : st.s11[1][1]) + (i > 0 ? st.s22[1][1] : 0.0);
    double* Hdi = Hd + i * BD;
    DBG_IDX(i, N, "diagonal block"); DBG_IDX(dvv, BD, "Hd entry"); DBG_IDX(dvx2, BD, "Hd entry");
    double vxx = fma(t00, qpq, cy.xx + uxx), vvv = fma(t11, qpq, cy.vv);
    if (!offd) {
      if (i == 0) { vxx += st.conf_prior_w; vvv += st.vel_prior_w; }
      if (last) { vxx += st.end_conf_prior_w; vvv += st.vel_prior_w; }
    }
    Hdi[dxx] = vxx;
    Hdi[dvx1] = fma(t10, qpq, cy.xv);
    if (offd) Hdi[dvx2] = fma(t10, qqp, cy.xv);
    Hdi[dvv] = vvv;
  }
  if (glane) {
    double* gi = g + i * b + lane;
    gi[0] += cy.gx + ugx;
    gi[D] += cy.gv;
  }
}


// ---- assembly of one PASS of the assembling linearize kernel (pk_linh_kernel): intervals [i0, i1), i1 - i0 <= 5, from
//      the 6 staged rows of each (stage[((i - i0) * 6 + j) * RS + entry]) into the pass buffer in shared memory
//      (HoP: i1 - i0 coupling blocks, HdP: diagonal blocks i0 .. i1 - 1, and block i1 = N - 1 in the last pass).
//      Same arithmetic, entry by entry, as assemble_intervals<D, 5> + store_block. ----
template <int D>
__device__ __forceinline__ void assemble_pass(const KSetting& st, const double* stage, double* HoP, double* HdP, double* g, double* carry,
                                              int N, int i0, int i1, int lane) {
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b, T = D * (D + 1) / 2, RS = (T + D + 1) & ~1, KS = 5;
  int p = (int)((sqrtf(8.0f * (float)lane + 1.0f) - 1.0f) * 0.5f);
  if (p * (p + 1) / 2 > lane) p--;
  if ((p + 1) * (p + 2) / 2 <= lane) p++;
  const int q = lane - p * (p + 1) / 2;
  const bool hlane = lane < T, glane = lane < D, offd = p != q;
  const int pc = hlane ? p : 0, qc = hlane ? q : 0;
  const double qpq = st.Qc_inv[pc * D + qc], qqp = st.Qc_inv[qc * D + pc];
  const int dxx = pc * (pc + 1) / 2 + qc, dvx1 = (D + pc) * (D + pc + 1) / 2 + qc, dvx2 = (D + qc) * (D + qc + 1) / 2 + pc,
            dvv = (D + pc) * (D + pc + 1) / 2 + D + qc;
  const int o1 = pc * b + qc, o2 = qc * b + pc;
  const int em = hlane ? lane : 0, eg = T + (glane ? lane : 0);
  // carry of the previous pass (this lane's three Hessian sums and two gradient sums)
  double cxx = 0.0, cxv = 0.0, cvv = 0.0, cgx = 0.0, cgv = 0.0;
  if (i0 > 0) {
    if (hlane) { cxx = carry[lane]; cxv = carry[T + lane]; cvv = carry[2 * T + lane]; }
    if (glane) { cgx = carry[3 * T + lane]; cgv = carry[3 * T + D + lane]; }
  }
#pragma unroll 1
  for (int i = i0; i < i1; i++) {
    const double* rows = stage + (size_t)(i - i0) * 6 * RS;
    double val[KS + 1], gval[KS + 1];
#pragma unroll
    for (int j = 0; j <= KS; j++) { val[j] = rows[j * RS + em]; gval[j] = rows[j * RS + eg]; }
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
      double* Hoi = HoP + (i - i0) * BB;
      Hoi[o1] = fma(st.s12[0][0], qpq, oxx);             Hoi[o1 + D] = fma(st.s12[0][1], qpq, oxv);
      Hoi[o1 + D * b] = fma(st.s12[1][0], qpq, ovx);     Hoi[o1 + D * b + D] = fma(st.s12[1][1], qpq, ovv);
      if (offd) {
        Hoi[o2] = fma(st.s12[0][0], qqp, oxx);           Hoi[o2 + D] = fma(st.s12[0][1], qqp, oxv);
        Hoi[o2 + D * b] = fma(st.s12[1][0], qqp, ovx);   Hoi[o2 + D * b + D] = fma(st.s12[1][1], qqp, ovv);
      }
      const double t00 = st.s11[0][0] + (i > 0 ? st.s22[0][0] : 0.0), t10 = st.s11[1][0] + (i > 0 ? st.s22[1][0] : 0.0),
                   t11 = st.s11[1][1] + (i > 0 ? st.s22[1][1] : 0.0);
      double* Hdi = HdP + (i - i0) * BD;
      double vxx = fma(t00, qpq, a0xx), vvv = fma(t11, qpq, a0vv);
      if (i == 0 && !offd) { vxx += st.conf_prior_w; vvv += st.vel_prior_w; }
      Hdi[dxx] = vxx;
      Hdi[dvx1] = fma(t10, qpq, a0xv);
      if (offd) Hdi[dvx2] = fma(t10, qqp, a0xv);
      Hdi[dvv] = vvv;
    }
    if (glane) {
      double* gi = g + i * b + lane;
      gi[0] += g0x;
      gi[D] += g0v;
    }
    cxx = a1xx; cxv = a1xv; cvv = a1vv; cgx = g1x; cgv = g1v;
  }
  if (i1 == N - 1) {
    // the last block: s22 + priors of x_T, v_T + carry + the unary factor of the last state (staged in row 30)
    const double uxx = stage[30 * RS + em], ugx = stage[30 * RS + eg];
    if (hlane) {
      const double t00 = st.s22[0][0], t10 = st.s22[1][0], t11 = st.s22[1][1];
      double* Hdi = HdP + (i1 - i0) * BD;
      double vxx = fma(t00, qpq, cxx + uxx), vvv = fma(t11, qpq, cvv);
      if (!offd) { vxx += st.end_conf_prior_w; vvv += st.vel_prior_w; }
      Hdi[dxx] = vxx;
      Hdi[dvx1] = fma(t10, qpq, cxv);
      if (offd) Hdi[dvx2] = fma(t10, qqp, cxv);
      Hdi[dvv] = vvv;
    }
    if (glane) {
      double* gi = g + (N - 1) * b + lane;
      gi[0] += cgx + ugx;
      gi[D] += cgv;
    }
  } else {
    if (hlane) { carry[lane] = cxx; carry[T + lane] = cxv; carry[2 * T + lane] = cvv; }
    if (glane) { carry[3 * T + lane] = cgx; carry[3 * T + D + lane] = cgv; }
  }
}

}  // namespace pkm

template <int D>
__global__ void __launch_bounds__(64, 8)
pk_solve_mma_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
                    const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b;
  const int N = st.N, K = st.K, NB = pk_even(N * b);
  double* g = smem;
  double* dl = g + NB;            // holds the states xs until the solve
  double* scr = dl + NB;          // 64 doubles: solver scratch; [0] doubles as the work-queue broadcast slot between solves
  double* Ho = scr + 64;
  double* Hd = Ho + (N - 1) * BB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, par = round & 1;
  const unsigned n = pr.pk_count[par * 2 + 1];
  const int32_t* list = pr.pk_lists + (size_t)(par * 2 + 1) * pr.B;
  const int SS = pk_state_size(D, N), RS = pk_row_stride(D);
  const size_t MLS = pk_mlist_size(D, N, K);
  mma::Solver<D> S;
  S.dbg_nb = N * b;
  unsigned long long n_solve = 0;
#ifdef PKM_PROFILE
  long long pt[6] = {0, 0, 0, 0, 0, 0}, tq;
#define PKM_PT(k) { const long long t1_ = clock64(); pt[k] += t1_ - tq; tq = t1_; }
#else
#define PKM_PT(k)
#endif
  // the two warps split the assembly the way they split the elimination: warp 0 owns blocks 0..m-1 (intervals [0, m)),
  // warp 1 blocks m..N-1 (intervals [m, N-1)) -- the right-state contribution of interval m-1 to block m is recomputed
  // by warp 1 (a few FMAs on six M-list rows), so neither warp ever waits for the other before its chain
  const int m = N / 2;
  // (tried: popping the work queue one trajectory ahead and prefetching the next trajectory's state and M-list into L2
  //  during the solve -- the assembly's DRAM round trips shrank, the solve slowed down by as much: 59.6 -> 61.5 ms)
  long long pos = blockIdx.x;
  while (pos < (long long)n) {
#ifdef PKM_PROFILE
    tq = clock64();
#endif
    const int64_t prob = list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    double* sp = pr.pk_state + prob * SS;
    double* sc = sp + 2 * pk_even(N * b);
    const double* ml = pr.pk_mlist + prob * MLS;
    for (int idx = tid; idx < N * b; idx += 64) dl[idx] = sp[idx];
    const double lambda = sc[PKS_LAMBDA];
    __syncthreads();
    PKM_PT(0)
    pkm::state_gradient<D>(st, dl, g, N, pr.start_conf + prob * D, pr.start_vel + prob * D, pr.end_conf + prob * D, pr.end_vel + prob * D, tid, 64);
    __syncthreads();
    PKM_PT(1)
    {
      const int ib = w ? m : 0, ie = w ? N - 1 : m;
      pkm::Carry cy; cy.xx = 0.0; cy.xv = 0.0; cy.vv = 0.0; cy.gx = 0.0; cy.gv = 0.0;
      if (K == 5) {
        if (w == 1 && m > 0) cy = pkm::right_state_carry<D, 5>(st, ml, RS, K, m - 1, lane);
        cy = pkm::assemble_intervals<D, 5>(st, ml, RS, Hd, Ho, g, N, K, ib, ie, cy, lane);
      } else {
        if (w == 1 && m > 0) cy = pkm::right_state_carry<D, -1>(st, ml, RS, K, m - 1, lane);
        cy = pkm::assemble_intervals<D, -1>(st, ml, RS, Hd, Ho, g, N, K, ib, ie, cy, lane);
      }
      if (w == 1) pkm::store_block<D>(st, ml, RS, Hd, g, N, K, N - 1, cy, lane);
      __syncwarp();
      pkm::limit_curvature<D>(st, dl, Hd, w ? m : 0, w ? N : m, lane, 32);
      __syncwarp();
    }
    PKM_PT(2)
    S.solve2(Hd, Ho, g, dl, lambda, N, scr, [] {});
    __syncthreads();
    PKM_PT(3)
    // linearized cost change = -(g.delta) - 0.5 delta^T H delta = -0.5 g.delta + 0.5 lambda |delta|^2
    double gd = 0.0, dd = 0.0;
    bool ok = true;
    double* dp = sp + pk_even(N * b);
    for (int idx = tid; idx < N * b; idx += 64) {
      const double dv = dl[idx];
      gd = fma(g[idx], dv, gd);
      dd = fma(dv, dv, dd);
      ok = ok && (fabs(dv) < CUDART_INF);
      dp[idx] = dv;
    }
    gd = warp_sum(gd); dd = warp_sum(dd);
    ok = __all_sync(FULL_MASK, ok);
    if (lane == 0) { scr[2 + 4 * w] = gd; scr[3 + 4 * w] = dd; scr[4 + 4 * w] = ok ? 1.0 : 0.0; }   // (the solver is done with scr)
    if (tid == 0) {
      const unsigned long long nx = atomicAdd(pr.queue, 1ull);
      scr[0] = (double)(long long)(nx + gridDim.x);
    }
    __syncthreads();
    if (tid == 0) {
      sc[PKS_LIN_COST_CHANGE] = -0.5 * (scr[2] + scr[6]) + 0.5 * lambda * (scr[3] + scr[7]);
      sc[PKS_SOLVED] = (scr[4] != 0.0 && scr[8] != 0.0) ? 1.0 : 0.0;
    }
    pos = (long long)scr[0];
    n_solve++;
    __syncthreads();                                   // dl, g and scr are free for the next trajectory
    PKM_PT(4)
  }
  if (tid == 0 && pr.counters && n_solve) atomicAdd(pr.counters + 1, n_solve);
#ifdef PKM_PROFILE
  if ((tid == 0 || tid == 32) && blockIdx.x == 5 && n_solve > 8)
    printf("pkm warp %d: %llu solves; per solve: load %lld, gradient %lld, assembly %lld, solve %lld, epilogue %lld clk\n", w, n_solve,
           pt[0] / (long long)n_solve, pt[1] / (long long)n_solve, pt[2] / (long long)n_solve, pt[3] / (long long)n_solve, pt[4] / (long long)n_solve);
#endif
}


// ------------------------------------------------------------------------------------------------------------------
// H-path (obs_check_inter = 5, the library default): the linearize kernel also ASSEMBLES.  Per pass of 5 intervals a
// warp evaluates the pass's 30 configurations (one per lane; the last state's unary factor in lane 30 of the last pass),
// stages their (M, cv) rows in shared memory, assembles the pass's coupling / diagonal blocks entry-parallel into a
// pass buffer and streams it to the trajectory's H in HBM with coalesced stores.  The solve kernel then only loads H:
// its assembly (24 k of its 64 k cycles per solve, mostly DRAM round trips on the M-list) is gone, and a rejected LM step
// -- 35 % of the solves -- re-solves with a new lambda without re-assembling anything.
// ------------------------------------------------------------------------------------------------------------------
template <class Opt>
__global__ void __launch_bounds__(32, PK_LIN_MIN_BLOCKS)
pk_linh_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
               const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim, b = 2 * D, BD = b * (b + 1) / 2, BB = b * b, T = D * (D + 1) / 2, RS = (T + D + 1) & ~1, IPP = PK_LINH_IPP;
  Opt o(rb, sdf, st, hconst, smem, false, 1);
  const int lane = o.lane, N = o.N, par = round & 1, NB = pk_even(N * b);
  double* xs = smem;
  double* g = xs + NB;
  double* stage = g + NB;
  double* carry = stage + 32 * RS;
  double* HoP = carry + pk_even(3 * T + 2 * D);
  double* HdP = HoP + IPP * BB;
  const unsigned n = pr.pk_count[par * 2 + 0];
  if (blockIdx.x == 0 && lane == 0) { pr.pk_count[(par ^ 1) * 2 + 0] = 0; pr.pk_count[(par ^ 1) * 2 + 1] = 0; }
  const int32_t* list = pk::list_of(pr, par, 0);
  const pk::Queue queue{pr.queue, lane};
  const int SS = pk_state_size(D, N);
  const size_t HS = pk_hbuf_size(D, N);
  const int HOFF = (N - 1) * BB, GOFF = pk_even((N - 1) * BB + N * BD);
  unsigned long long n_lin = 0;
  for (int64_t pos = blockIdx.x; pos < (int64_t)n; pos = queue.next()) {
    const int64_t prob = list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    // sphere masks of a pass (one per lane = configuration), loaded one pass ahead so that the load overlaps other work
    auto mask_of = [&](int p0) -> unsigned long long {
      const int p1 = min(p0 + IPP, N - 1), pn = p1 - p0;
      const bool pcfg = lane < 6 * pn, ptail = p1 == N - 1 && lane == 30;
      const int cidx = ptail ? o.C - 1 : (p0 + lane / 6) * 6 + lane % 6;
      if (pcfg || ptail) DBG_IDX(cidx, o.C, "sphere-mask configuration index");
      return (pcfg || ptail) ? pr.pk_mask[prob * o.C + cidx] : 0ull;
    };
    unsigned long long sm_next = pr.pk_mask_use ? mask_of(0) : 0ull;
    pk::load_states(o, pr.pk_state + prob * SS, false);
    __syncwarp();
    pkm::state_gradient<D>(st, xs, g, N, pr.start_conf + prob * D, pr.start_vel + prob * D, pr.end_conf + prob * D, pr.end_vel + prob * D, lane, 32);
    __syncwarp();
    double* H = pr.pk_mlist + prob * HS;
#pragma unroll 1
    for (int i0 = 0; i0 < N - 1; i0 += IPP) {
      const int i1 = min(i0 + IPP, N - 1), ni = i1 - i0;
      const bool last = i1 == N - 1;
      // ---- configuration pass: lane l < 6 ni <-> (interval i0 + l / 6, configuration l % 6); lane 30: the last state ----
      {
        double M[T], cv[D];
#pragma unroll
        for (int m = 0; m < T; m++) M[m] = 0.0;
#pragma unroll
        for (int d = 0; d < D; d++) cv[d] = 0.0;
        const bool cfg = lane < 6 * ni, tail = last && lane == 30;
        const int i = tail ? N - 1 : i0 + lane / 6, j = tail ? 0 : lane % 6;
        if (pr.pk_mask_use) {
          // spheres within reach of their hinge, recorded by the error evaluation of these very states
          const unsigned long long sm = sm_next;
          if (i1 < N - 1) sm_next = mask_of(i1);
          const unsigned lo = __reduce_or_sync(FULL_MASK, (unsigned)sm), hi = __reduce_or_sync(FULL_MASK, (unsigned)(sm >> 32));
          const unsigned long long wm = ((unsigned long long)hi << 32) | lo;
          if (wm != 0ull && (cfg || tail)) {
            double e2 = 0.0, es = 0.0;
            config_eval<D, Opt::NDim, 0, true, false, true>(rb, sdf, o.template config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, M, cv,
                                                            e2, es, nullptr, nullptr, sm, wm);
          }
        } else if (cfg || tail) {
          double e2 = 0.0, es = 0.0;
          config_eval<D, Opt::NDim, 0, true, false>(rb, sdf, o.template config_state<false>(i, j), st.epsilon, st.inv_cost_sigma, M, cv, e2, es,
                                                    nullptr, nullptr);
        }
        double* row = stage + lane * RS;
#pragma unroll
        for (int m = 0; m < T; m++) row[m] = M[m];
#pragma unroll
        for (int d = 0; d < D; d++) row[T + d] = cv[d];
      }
      __syncwarp();
      // ---- entry-parallel assembly of the pass, limit curvature, coalesced flush ----
      pkm::assemble_pass<D>(st, stage, HoP, HdP, g, carry, N, i0, i1, lane);
      __syncwarp();
      if (st.flag_pos_limit || st.flag_vel_limit) {
        pkm::limit_curvature<D>(st, xs + i0 * b, HdP, 0, ni + (last ? 1 : 0), lane, 32);
        __syncwarp();
      }
      {
        const double2* src = reinterpret_cast<const double2*>(HoP);
        double2* dst = reinterpret_cast<double2*>(H + (size_t)i0 * BB);
        for (int idx = lane; idx < ni * BB / 2; idx += 32) dst[idx] = src[idx];
        const int nd = (ni + (last ? 1 : 0)) * BD;
        const double* s1 = HdP;
        double* d1 = H + HOFF + (size_t)i0 * BD;
        for (int idx = lane; idx < nd; idx += 32) d1[idx] = s1[idx];
      }
      __syncwarp();
    }
    for (int idx = lane; idx < N * b; idx += 32) H[GOFF + idx] = g[idx];
    n_lin++;
    __syncwarp();
  }
  if (lane == 0 && pr.counters && n_lin) atomicAdd(pr.counters + 0, n_lin);
}

template <int D>
__global__ void __launch_bounds__(64, 8)
pk_solve_mma_h_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
                      const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b;
  const int N = st.N, NB = pk_even(N * b);
  double* g = smem;
  double* dl = g + NB;
  double* scr = dl + NB;
  double* Ho = scr + 64;
  double* Hd = Ho + (N - 1) * BB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, par = round & 1;
  const unsigned n = pr.pk_count[par * 2 + 1];
  const int32_t* list = pr.pk_lists + (size_t)(par * 2 + 1) * pr.B;
  const int SS = pk_state_size(D, N);
  const size_t HS = pk_hbuf_size(D, N);
  const int HSZ = (N - 1) * BB + N * BD, GOFF = pk_even(HSZ);
  mma::Solver<D> S;
  S.dbg_nb = N * b;
  unsigned long long n_solve = 0;
  long long pos = blockIdx.x;
  while (pos < (long long)n) {
    const int64_t prob = list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    double* sp = pr.pk_state + prob * SS;
    double* sc = sp + 2 * pk_even(N * b);
    const double* H = pr.pk_mlist + prob * HS;
    // H (Ho | Hd, contiguous in both places) and g: 16-byte asynchronous copies straight into shared memory
    {
      const unsigned sH = (unsigned)__cvta_generic_to_shared(Ho), sG = (unsigned)__cvta_generic_to_shared(g);
      for (int idx = tid; idx < (HSZ + 1) / 2; idx += 64)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sH + 16u * idx), "l"(H + 2 * idx) : "memory");
      for (int idx = tid; idx < (N * b) / 2; idx += 64)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sG + 16u * idx), "l"(H + GOFF + 2 * idx) : "memory");
    }
    const double lambda = sc[PKS_LAMBDA];
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    S.solve2(Hd, Ho, g, dl, lambda, N, scr, [] {});
    __syncthreads();
    double gd = 0.0, dd = 0.0;
    bool ok = true;
    double* dp = sp + pk_even(N * b);
    for (int idx = tid; idx < N * b; idx += 64) {
      const double dv = dl[idx];
      gd = fma(g[idx], dv, gd);
      dd = fma(dv, dv, dd);
      ok = ok && (fabs(dv) < CUDART_INF);
      dp[idx] = dv;
    }
    gd = warp_sum(gd); dd = warp_sum(dd);
    ok = __all_sync(FULL_MASK, ok);
    if (lane == 0) { scr[2 + 4 * w] = gd; scr[3 + 4 * w] = dd; scr[4 + 4 * w] = ok ? 1.0 : 0.0; }
    if (tid == 0) {
      const unsigned long long nx = atomicAdd(pr.queue, 1ull);
      scr[0] = (double)(long long)(nx + gridDim.x);
    }
    __syncthreads();
    if (tid == 0) {
      sc[PKS_LIN_COST_CHANGE] = -0.5 * (scr[2] + scr[6]) + 0.5 * lambda * (scr[3] + scr[7]);
      sc[PKS_SOLVED] = (scr[4] != 0.0 && scr[8] != 0.0) ? 1.0 : 0.0;
    }
    pos = (long long)scr[0];
    n_solve++;
    __syncthreads();
  }
  if (tid == 0 && pr.counters && n_solve) atomicAdd(pr.counters + 1, n_solve);
}
