// mma_solve.cuh -- damped block-tridiagonal Cholesky solve of one trajectory's normal equations on the FP64 tensor
// cores (DMMA, mma.sync.m8n8k4.f64), one warp per trajectory.
//
//   (H + lambda I) delta = -g,   H block tridiagonal with N diagonal blocks of b x b (b = 2 D <= 14)
//
// replaces what GTSAM's sparse elimination does for the reference (gpmp2::optimize ->
// LevenbergMarquardtOptimizer::tryLambda -> GaussianFactorGraph::optimize, gpmp2/planner/BatchTrajOptimizer.cpp:212-308,
// SURVEY.md App. B.1) for the chain-structured graphs of internal::BatchTrajOptimize (BatchTrajOptimizer-inl.h:19-84).
//
// Why tensor cores for 14 x 14 blocks: the scalar solver of optimizer_kernel.cuh spends ~27 k warp instructions per
// solve on the panel factorizations and Schur updates, every DFMA fed by a shared-memory broadcast (LSU pipe 76 % busy,
// FP64 pipe 20 %).  A DMMA does 256 FMAs from two operand registers per lane, and on B200 it runs at the DFMA peak
// (37 TFLOP/s measured, scripts/micro/dmma_bench.cu), so the O(b^3) part of the factorization costs ~25 instructions
// per block instead of ~2000.
//
// Scheme (validated at matrix level by scripts/proto/mma_solve_proto.py):
//  * every block is padded to P = 8 NT (NT = 2 for b >= 8) and lives in registers as 8 x 8 accumulator tiles.  The
//    tile columns are stored PERMUTED -- accumulator register e0 of lane (r = lane / 4, j = lane % 4) is logical
//    column j, e1 is logical column 4 + j -- so that a 4-column slab of a tile already is the A fragment of the next
//    DMMA: no layout conversion between an accumulator and an operand, ever;
//  * right-looking factorization in panels of 4 columns (= the k of the DMMA).  Per panel: the 4 x 4 diagonal block
//    is gathered through 16 doubles of shared memory and factored + inverted redundantly by every lane (two
//    reciprocal square roots deep, the pivots are taken in pairs through the 2 x 2 determinant); ONE DMMA per
//    8-row group multiplies the panel by W^T = L44^-T -- its B operand is laid out so that e0 of the result is the
//    finished panel as an A fragment and e1 its negative; the same fragment, rows permuted by one shuffle, is the B
//    operand of the rank-4 updates of everything to the right (the rest of the block D, the coupling block C, the
//    Schur complement N for the next block): one DMMA per 8 x 8 tile;
//  * the right-hand side -g rides along as padding row b of the diagonal tiles: the panel solves forward-substitute it
//    for free, and putting that row into the A operand of the Schur update carries it into the next block's row b;
//  * two elimination chains per warp in one instruction stream (blocks 0.. downwards, N-1.. upwards, meeting at the
//    middle block, as the scalar solver does): twice the independent work per warp, and the redundant 4 x 4
//    factorizations of the two chains run in the two half-warps at once;
//  * the factor overwrites H in shared memory (L_ii packed lower with 1 / l_rr on the diagonal, Z_i row-major), the
//    back substitution runs both chains outwards from the middle in the two half-warps.
//  A non-positive pivot is not tested: its rsqrt is NaN, which reaches delta -- the caller tests delta once.
#pragma once
#include "kparams.h"

namespace mma {

struct Tile { double e0, e1; };

__device__ __forceinline__ void dmma(Tile& c, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.e0), "+d"(c.e1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double rsqrt64(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RSQ64H, ~2^-22 relative
  // one third-order step: y (1 + r/2 + 3 r^2/8), r = 1 - x y^2  ->  relative error ~2^-65
  const double r = fma(-x * y, y, 1.0);
  return fma(y * r, fma(0.375, r, 0.5), y);
}

template <int D>
struct Solver {
  static constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b;
  static constexpr int NT = (b + 8) / 8;         // tiles per block side: b real rows + the rhs row
  static constexpr int NP = (b + 3) / 4;         // panels holding real columns
  static constexpr int IR = b >> 3, RR = b & 7;  // tile row / local row of the rhs row
  static constexpr int NTC = (b + 7) / 8;        // tile rows / columns that hold real rows / columns
  static constexpr int SCR = 32;                 // scratch doubles per chain: gathered 4 x 4 block | W
  static constexpr unsigned FULL = 0xffffffffu;

  struct Chain {
    Tile Dt[NT][NT];     // this block (lower tiles I >= J; diagonal tiles full and symmetric), row b = rhs
    Tile Ct[NT][NT];     // coupling block: rows = the next block of the sweep, columns = this block
    Tile Nt[NT][NT];     // Schur complement accumulated for the next block (lower tiles), row b = its rhs update
    Tile Et[NTC][NTC];   // (look-ahead path) identity rows riding through the panels: end up as L_ii^-T (upper tiles I <= J)
    double *L, *Z, *y;   // where this block's factor goes: L_ii (packed lower), Z (row-major b x b), y (b)
  };

  int lane, r, j, srcB;
  int dbg_nb = 1 << 30;   // (GPMP2B_DEBUG_BOUNDS: length of the delta array, set by the caller)
#ifdef MMA_SOLVE_PROFILE
  mutable long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // cycles: chain forward, wait, middle, wait, back substitution
#define MMA_PT(k, t0) pt[k] += clock64() - t0
#else
#define MMA_PT(k, t0)
#endif

  __device__ __forceinline__ Solver() {
    lane = threadIdx.x & 31;
    r = lane >> 2; j = lane & 3;
    const int n = lane >> 2;                         // physical column of a B fragment held by this lane
    srcB = 4 * ((n >> 1) + 4 * (n & 1)) + j;         // lane holding row pi^-1(n) of the same fragment as an A operand
  }

  // ---- tile loads from the shared-memory H: branch-free (clamped offsets + selects; divergent branches around the
  //      loads cost a third of the solve in the first version) ----
  // element (R, c) of the block augmented by the rhs row: packed lower H for R < b, g for R == b, 0 otherwise
  __device__ __forceinline__ double d_elem(const double* Hd_i, const double* g_i, int R, int c, bool diag_tile) const {
    const bool ok = (R <= b) && (c < b);
    const int cc = min(c, b - 1), Rc = min(R, b - 1);
    const int hi = diag_tile ? max(Rc, cc) : Rc, lo = diag_tile ? min(Rc, cc) : cc;
    const double* p = (R == b) ? (g_i + cc) : (Hd_i + (hi * (hi + 1) / 2 + lo));
    const double v = *p;
    return ok ? v : 0.0;
  }
  __device__ __forceinline__ void load_D(Chain& ch, const double* Hd_i, const double* g_i, double lam0, double lam1) const {
#pragma unroll
    for (int I = 0; I < NT; I++)
#pragma unroll
      for (int J = 0; J <= I; J++) {
        if (8 * J >= b) { ch.Dt[I][J].e0 = 0.0; ch.Dt[I][J].e1 = 0.0; continue; }
        double v0 = d_elem(Hd_i, g_i, 8 * I + r, 8 * J + j, I == J) + ch.Nt[I][J].e0;
        double v1 = d_elem(Hd_i, g_i, 8 * I + r, 8 * J + 4 + j, I == J) + ch.Nt[I][J].e1;
        if (I == J) { v0 += lam0; v1 += lam1; }        // lambda on the diagonal (lam0 / lam1: this lane's share)
        ch.Dt[I][J].e0 = v0; ch.Dt[I][J].e1 = v1;
        ch.Nt[I][J].e0 = 0.0; ch.Nt[I][J].e1 = 0.0;
      }
  }
  // C[R][c] = Ho[R * sR + c * sC]  (down chain: Ho_i transposed, sR = 1, sC = b; up chain: Ho_{i-1}, sR = b, sC = 1)
  __device__ __forceinline__ void load_C(Chain& ch, const double* Ho_i, int sR, int sC) const {
#pragma unroll
    for (int I = 0; I < NT; I++)
#pragma unroll
      for (int J = 0; J < NT; J++) {
        double v0 = 0.0, v1 = 0.0;
        if (8 * I < b && 8 * J < b) {
          const int R = 8 * I + r, c0 = 8 * J + j, c1 = c0 + 4;
          const int Rc = min(R, b - 1);
          const double t0 = Ho_i[Rc * sR + min(c0, b - 1) * sC], t1 = Ho_i[Rc * sR + min(c1, b - 1) * sC];
          v0 = (R < b && c0 < b) ? t0 : 0.0;
          v1 = (R < b && c1 < b) ? t1 : 0.0;
        }
        ch.Ct[I][J].e0 = v0; ch.Ct[I][J].e1 = v1;
      }
  }
  __device__ __forceinline__ void zero_N(Chain& ch) const {
#pragma unroll
    for (int I = 0; I < NT; I++)
#pragma unroll
      for (int J = 0; J < NT; J++) { ch.Nt[I][J].e0 = 0.0; ch.Nt[I][J].e1 = 0.0; }
  }

  // ---- one panel (4 columns) of NCH chains ----
  template <int NCH, bool HASC, int p>
  __device__ __forceinline__ void panel(Chain (&ch)[NCH], double* scr) const {
    constexpr int J = p >> 1, h = p & 1;
    constexpr bool PAD2 = (b - 4 * p) == 2;          // columns 2, 3 of this panel are padding: identity
    const int hw = (NCH == 2) ? (lane >> 4) : 0;
    // 1. the 4 x 4 diagonal block -> scratch
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const double a = h ? ch[c].Dt[J][J].e1 : ch[c].Dt[J][J].e0;
      if ((r >> 2) == h) scr[c * SCR + (r & 3) * 4 + j] = a;
    }
    __syncwarp();
    // 2. L44 = chol(A44), W = L44^-1 (lower), redundantly in every lane of the chain's half-warp
    const double* gs = scr + hw * SCR;
    double w00, w10, w11, w20 = 0.0, w21 = 0.0, w22 = 1.0, w30 = 0.0, w31 = 0.0, w32 = 0.0, w33 = 1.0;
    {
      const double a00 = gs[0];
      const double2 a1 = *reinterpret_cast<const double2*>(gs + 4);     // a10 a11
      const double i0 = rsqrt64(a00);
      const double det = fma(a00, a1.y, -a1.x * a1.x);                  // l11^2 = det / a00
      const double i1 = rsqrt64(det) * (a00 * i0);
      const double l10 = a1.x * i0;
      w00 = i0; w11 = i1;
      w10 = -l10 * i0 * i1;
      if (!PAD2) {
        const double2 a2 = *reinterpret_cast<const double2*>(gs + 8);   // a20 a21
        const double a22 = gs[10];
        const double2 a3 = *reinterpret_cast<const double2*>(gs + 12);  // a30 a31
        const double2 a3b = *reinterpret_cast<const double2*>(gs + 14); // a32 a33
        const double l20 = a2.x * i0, l30 = a3.x * i0;
        const double l21 = fma(-l20, l10, a2.y) * i1, l31 = fma(-l30, l10, a3.y) * i1;
        const double b22 = fma(-l21, l21, fma(-l20, l20, a22));
        const double b32 = fma(-l31, l21, fma(-l30, l20, a3b.x));
        const double b33 = fma(-l31, l31, fma(-l30, l30, a3b.y));
        const double i2 = rsqrt64(b22);
        const double det2 = fma(b22, b33, -b32 * b32);
        const double i3 = rsqrt64(det2) * (b22 * i2);
        const double l32 = b32 * i2;
        w22 = i2; w33 = i3;
        w21 = -l21 * i1 * i2;
        w32 = -l32 * i2 * i3;
        w20 = -fma(l21, w10, l20 * i0) * i2;
        w31 = -fma(l32, w21, l31 * i1) * i3;
        w30 = -fma(l32, w20, fma(l31, w10, l30 * i0)) * i3;
      }
    }
    // 3. W -> scratch (row-major 4 x 4; the entries above the diagonal were zeroed once by the caller)
    if ((lane & 15) == 0 || (NCH == 1 && false)) {
      double* ws = scr + hw * SCR + 16;
      *reinterpret_cast<double2*>(ws + 0) = make_double2(w00, 0.0);
      *reinterpret_cast<double2*>(ws + 4) = make_double2(w10, w11);
      *reinterpret_cast<double2*>(ws + 8) = make_double2(w20, w21);
      *reinterpret_cast<double2*>(ws + 10) = make_double2(w22, 0.0);
      *reinterpret_cast<double2*>(ws + 12) = make_double2(w30, w31);
      *reinterpret_cast<double2*>(ws + 14) = make_double2(w32, w33);
    }
    __syncwarp();
    // B operand of the panel solve: physical column n = lane / 4 of the result is +f[.][n / 2] (n even), -f (n odd)
    double wb[NCH], wdg[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const double raw = scr[c * SCR + 16 + (lane >> 3) * 4 + j];
      wb[c] = ((lane >> 2) & 1) ? -raw : raw;
      wdg[c] = scr[c * SCR + 16 + j * 5];            // W[j][j] = 1 / l_jj of this lane's column
    }
    // 4. panel solves, 5. rank-4 updates
    const bool keep = h ? (r >= 4 && j <= r - 4) : (r >= 4 || j <= r);
    const int col = 4 * p + j;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      Tile fD[NT], fC[NT];
#pragma unroll
      for (int I = J; I < NT; I++) {
        Tile f; f.e0 = 0.0; f.e1 = 0.0;
        dmma(f, h ? ch[c].Dt[I][J].e1 : ch[c].Dt[I][J].e0, wb[c]);
        if (I == J && !keep) { f.e0 = 0.0; f.e1 = 0.0; }
        fD[I] = f;
        const int R = 8 * I + r;
        if (8 * I < b && R < b && col <= R) { DBG_IDX(R * (R + 1) / 2 + col, BD, "L"); ch[c].L[R * (R + 1) / 2 + col] = (R == col) ? wdg[c] : f.e0; }
        if (I == IR && r == RR && col < b) ch[c].y[col] = f.e0;
      }
      if (HASC) {
#pragma unroll
        for (int I = 0; I < NTC; I++) {
          Tile f; f.e0 = 0.0; f.e1 = 0.0;
          dmma(f, h ? ch[c].Ct[I][J].e1 : ch[c].Ct[I][J].e0, wb[c]);
          fC[I] = f;
          const int R = 8 * I + r;
          if (R < b && col < b) { DBG_IDX(R * b + col, BB, "Z"); ch[c].Z[R * b + col] = f.e0; }
        }
      }
      // the rest of this block and of the coupling block
#pragma unroll
      for (int Jt = (h ? J + 1 : J); Jt < NTC; Jt++) {
        double bD = __shfl_sync(FULL, fD[Jt].e0, srcB);
        if (8 * Jt + 7 >= b) {                                          // the rhs row and padding rows never act as columns
          const int n = lane >> 2;
          if (8 * Jt + (n >> 1) + 4 * (n & 1) >= b) bD = 0.0;
        }
#pragma unroll
        for (int I = Jt; I < NT; I++) dmma(ch[c].Dt[I][Jt], fD[I].e1, bD);
        if (HASC) {
#pragma unroll
          for (int I = 0; I < NTC; I++) dmma(ch[c].Ct[I][Jt], fC[I].e1, bD);
        }
      }
      // Schur complement for the next block; the rhs row of this block rides in the A operand
      if (HASC) {
        double bC[NTC];
#pragma unroll
        for (int Jn = 0; Jn < NTC; Jn++) bC[Jn] = __shfl_sync(FULL, fC[Jn].e0, srcB);
#pragma unroll
        for (int I = 0; I < NT; I++) {
          double A = (I < NTC) ? fC[I < NTC ? I : 0].e1 : 0.0;
          if (I == IR && r == RR) A = fD[IR].e1;
#pragma unroll
          for (int Jn = 0; Jn <= I && Jn < NTC; Jn++) dmma(ch[c].Nt[I][Jn], A, bC[Jn]);
        }
      }
    }
  }

  template <int NCH, bool HASC, int p>
  __device__ __forceinline__ void panels(Chain (&ch)[NCH], double* scr) const {
    if constexpr (p < NP) {
      panel<NCH, HASC, p>(ch, scr);
      panels<NCH, HASC, p + 1>(ch, scr);
    }
  }

  // ---- the whole solve.  Hd: N packed-lower diagonal blocks, Ho: N - 1 row-major blocks H_{i,i+1}, g: gradient,
  //      dl: delta out; scr: 2 * SCR doubles, 16-byte aligned; H is overwritten by the factor. ----
  __device__ void solve(double* Hd, double* Ho, const double* g, double* dl, double lambda, int N, double* scr) const {
    const int m = N / 2, nd = m, nu = N - 1 - m;     // down chain: blocks 0..m-1, up chain: N-1..m+1
    for (int idx = lane; idx < 2 * SCR; idx += 32) scr[idx] = 0.0;
    Chain ch[2];
    zero_N(ch[0]); zero_N(ch[1]);
    // the rhs row carries +g (so x = (H + lambda I)^-1 g comes out); delta = -x is written by the back substitution
    const double lam0 = (r == j) ? lambda : 0.0, lam1 = (r == j + 4) ? lambda : 0.0;
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < nu; s++) {
      const int i0 = s, i1 = N - 1 - s;
      load_D(ch[0], Hd + i0 * BD, g + i0 * b, lam0, lam1);
      load_C(ch[0], Ho + i0 * BB, 1, b);
      load_D(ch[1], Hd + i1 * BD, g + i1 * b, lam0, lam1);
      load_C(ch[1], Ho + (i1 - 1) * BB, b, 1);
      ch[0].L = Hd + i0 * BD; ch[0].Z = Ho + i0 * BB; ch[0].y = dl + i0 * b;
      ch[1].L = Hd + i1 * BD; ch[1].Z = Ho + (i1 - 1) * BB; ch[1].y = dl + i1 * b;
      __syncwarp();
      panels<2, true, 0>(ch, scr);
    }
    if (nd > nu) {   // even N: one more block of the down chain
      Chain (&c1)[1] = reinterpret_cast<Chain (&)[1]>(ch[0]);
      const int i0 = nu;
      load_D(ch[0], Hd + i0 * BD, g + i0 * b, lam0, lam1);
      load_C(ch[0], Ho + i0 * BB, 1, b);
      ch[0].L = Hd + i0 * BD; ch[0].Z = Ho + i0 * BB; ch[0].y = dl + i0 * b;
      __syncwarp();
      panels<1, true, 0>(c1, scr);
    }
    {                // the middle block takes both Schur complements
      Chain (&c1)[1] = reinterpret_cast<Chain (&)[1]>(ch[0]);
#pragma unroll
      for (int I = 0; I < NT; I++)
#pragma unroll
        for (int J = 0; J <= I; J++) { ch[0].Nt[I][J].e0 += ch[1].Nt[I][J].e0; ch[0].Nt[I][J].e1 += ch[1].Nt[I][J].e1; }
      load_D(ch[0], Hd + m * BD, g + m * b, lam0, lam1);
      ch[0].L = Hd + m * BD; ch[0].Z = nullptr; ch[0].y = dl + m * b;
      __syncwarp();
      panels<1, false, 0>(c1, scr);
    }
    __syncwarp();
    // ---- back substitution: half-warp 0 walks the down chain back (m-1 .. 0), half-warp 1 the up chain (m+1 .. N-1) ----
    const int hw = lane >> 4, c = lane & 15;
    const bool act = c < b;
    const int nsteps = max(nd, nu);
#pragma unroll 1
    for (int s = 0; s <= nsteps; s++) {
      const int i = s == 0 ? m : (hw ? m + s : m - s);
      const bool valid = act && (s == 0 || (hw ? s <= nu : s <= nd));
      const int ic = valid ? i : m;
      double t = valid ? dl[ic * b + c] : 0.0;          // y_i (forward-substituted +g)
      if (s > 0) {                                      // t = y_i - Z_i^T x_prev = y_i + Z_i^T delta_prev  (branch-free)
        const double* Zp = Ho + (valid ? (hw ? ic - 1 : ic) : 0) * BB + min(c, b - 1);
        const double* xp = dl + (valid ? (hw ? ic - 1 : ic + 1) : 0) * b;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int R = 0; R < b; R += 2) {
          const double2 x2 = *reinterpret_cast<const double2*>(xp + R);
          t0 = fma(Zp[R * b], x2.x, t0);
          t1 = fma(Zp[(R + 1) * b], x2.y, t1);
        }
        t = valid ? t + (t0 + t1) : 0.0;
      }
      // L^T x = t, column-oriented, pre-scaled by 1 / l_cc so that a step is SHFL -> DFMA
      const double* Lp = Hd + ic * BD;
      const int cl = min(c, b - 1);
      const double idl = Lp[cl * (cl + 1) / 2 + cl];
      const double idc = valid ? idl : 0.0;
      double Lc[b];
#pragma unroll
      for (int R = 0; R < b; R++) { const double l = Lp[R * (R + 1) / 2 + min(cl, R)]; Lc[R] = (R > c ? l : 0.0) * idc; }
      double u = t * idc;
#pragma unroll
      for (int R = b - 1; R >= 1; R--) {
        const double xr = __shfl_sync(FULL, u, (lane & 16) | R);
        u = fma(-Lc[R], xr, u);
      }
      __syncwarp();
      if (valid && (s > 0 || hw == 0)) dl[ic * b + c] = -u;
      __syncwarp();
    }
  }

  // =====================================================================================================
  // Single chain per warp with LOOK-AHEAD (used by solve2).  The dependent chain of a block is
  //   4 x 4 factorization of panel p -> panel solve of the tile row that holds the next diagonal block -> its rank-4
  //   update -> gather -> 4 x 4 factorization of panel p + 1,
  // everything else of panel p (the other panel solves, the stores of the factor, the updates of the coupling block
  // and of the Schur complement) is "bulk" work that only has to be issued at some point: it is placed in the same
  // basic block as the NEXT panel's 4 x 4 factorization, so the scheduler overlaps it with that latency chain.
  // =====================================================================================================
  struct PanelState { Tile fD[NT], fC[NT], fE[NTC]; double wb, wdg, bDX; };

  template <int p>
  __device__ __forceinline__ void gather(const Chain& ch, double* scr) const {
    constexpr int J = p >> 1, h = p & 1;
    const double a = h ? ch.Dt[J][J].e1 : ch.Dt[J][J].e0;
    if ((r >> 2) == h) { DBG_IDX((r & 3) * 4 + j, 16, "gather scratch"); scr[(r & 3) * 4 + j] = a; }
  }

  // L44 = chol(A44), W = L44^-1 (lower) from the gathered block, redundantly in every lane; W -> scratch
  template <int p>
  __device__ __forceinline__ void potrf4(double* scr) const {
    constexpr bool PAD2 = (b - 4 * p) == 2;          // columns 2, 3 of this panel are padding: identity
    const double* gs = scr;
    double w00, w10, w11, w20 = 0.0, w21 = 0.0, w22 = 1.0, w30 = 0.0, w31 = 0.0, w32 = 0.0, w33 = 1.0;
    const double a00 = gs[0];
    const double2 a1 = *reinterpret_cast<const double2*>(gs + 4);     // a10 a11
    const double i0 = rsqrt64(a00);
    const double det = fma(a00, a1.y, -a1.x * a1.x);                  // l11^2 = det / a00
    const double i1 = rsqrt64(det) * (a00 * i0);
    const double l10 = a1.x * i0;
    w00 = i0; w11 = i1;
    w10 = -l10 * i0 * i1;
    if (!PAD2) {
      const double2 a2 = *reinterpret_cast<const double2*>(gs + 8);   // a20 a21
      const double a22 = gs[10];
      const double2 a3 = *reinterpret_cast<const double2*>(gs + 12);  // a30 a31
      const double2 a3b = *reinterpret_cast<const double2*>(gs + 14); // a32 a33
      const double l20 = a2.x * i0, l30 = a3.x * i0;
      const double l21 = fma(-l20, l10, a2.y) * i1, l31 = fma(-l30, l10, a3.y) * i1;
      const double b22 = fma(-l21, l21, fma(-l20, l20, a22));
      const double b32 = fma(-l31, l21, fma(-l30, l20, a3b.x));
      const double b33 = fma(-l31, l31, fma(-l30, l30, a3b.y));
      const double i2 = rsqrt64(b22);
      const double det2 = fma(b22, b33, -b32 * b32);
      const double i3 = rsqrt64(det2) * (b22 * i2);
      const double l32 = b32 * i2;
      w22 = i2; w33 = i3;
      w21 = -l21 * i1 * i2;
      w32 = -l32 * i2 * i3;
      w20 = -fma(l21, w10, l20 * i0) * i2;
      w31 = -fma(l32, w21, l31 * i1) * i3;
      w30 = -fma(l32, w20, fma(l31, w10, l30 * i0)) * i3;
    }
    if (lane == 0) {
      double* ws = scr + 16;
      *reinterpret_cast<double2*>(ws + 0) = make_double2(w00, 0.0);
      *reinterpret_cast<double2*>(ws + 4) = make_double2(w10, w11);
      *reinterpret_cast<double2*>(ws + 8) = make_double2(w20, w21);
      *reinterpret_cast<double2*>(ws + 10) = make_double2(w22, 0.0);
      *reinterpret_cast<double2*>(ws + 12) = make_double2(w30, w31);
      *reinterpret_cast<double2*>(ws + 14) = make_double2(w32, w33);
    }
  }

  __device__ __forceinline__ double bfrag(double f, int Jt) const {
    double bD = __shfl_sync(FULL, f, srcB);
    if (8 * Jt + 7 >= b) {                             // the rhs row and padding rows never act as columns
      const int n = lane >> 2;
      if (8 * Jt + (n >> 1) + 4 * (n & 1) >= b) bD = 0.0;
    }
    return bD;
  }
  template <int p>
  __device__ __forceinline__ Tile panel_solve(const Tile& t, double wb, bool is_diag_tile) const {
    constexpr int h = p & 1;
    Tile f; f.e0 = 0.0; f.e1 = 0.0;
    dmma(f, h ? t.e1 : t.e0, wb);
    if (is_diag_tile) {
      const bool keep = h ? (r >= 4 && j <= r - 4) : (r >= 4 || j <= r);
      if (!keep) { f.e0 = 0.0; f.e1 = 0.0; }
    }
    return f;
  }

  // the dependent part of panel p: up to the gather of panel p + 1
  template <int p>
  __device__ __forceinline__ void crit(Chain& ch, PanelState& ps, double* scr) const {
    constexpr int J = p >> 1, h = p & 1, X = h ? J + 1 : J;
    const double raw = scr[16 + (lane >> 3) * 4 + j];
    ps.wb = ((lane >> 2) & 1) ? -raw : raw;          // B operand of the panel solve: +f in e0, -f in e1
    ps.wdg = scr[16 + j * 5];                         // W[j][j] = 1 / l_jj of this lane's column
    if constexpr (p + 1 < NP) {
      ps.fD[X] = panel_solve<p>(ch.Dt[X][J], ps.wb, X == J);
      ps.bDX = bfrag(ps.fD[X].e0, X);
      dmma(ch.Dt[X][X], ps.fD[X].e1, ps.bDX);
      gather<p + 1>(ch, scr);
    }
  }

  // everything else of panel p
  template <bool HASC, int p>
  __device__ __forceinline__ void bulk(Chain& ch, PanelState& ps) const {
    constexpr int J = p >> 1, h = p & 1, X = h ? J + 1 : J;
    constexpr bool HAVE_X = (p + 1 < NP);
    const int col = 4 * p + j;
#pragma unroll
    for (int I = J; I < NT; I++) {
      if (!(HAVE_X && I == X)) ps.fD[I] = panel_solve<p>(ch.Dt[I][J], ps.wb, I == J);
      if (I == IR && r == RR && col < b) ch.y[col] = ps.fD[I].e0;
    }
    // identity rows: what the panel operations make of them is L_ii^-T, which is all the back substitution needs of the
    // diagonal block (stored packed by columns, U[R][col] at col (col + 1) / 2 + R, in the block's Hd slot; L itself is
    // never stored).  Row R of E is zero left of column R, so only tile rows I <= J take part, unmasked.
#pragma unroll
    for (int I = 0; I <= J && I < NTC; I++) {
      ps.fE[I] = panel_solve<p>(ch.Et[I][J], ps.wb, false);
      const int R = 8 * I + r;
      if (col < b && R <= col) { DBG_IDX(col * (col + 1) / 2 + R, BD, "L^-T"); ch.L[col * (col + 1) / 2 + R] = ps.fE[I].e0; }
    }
    if (HASC) {
#pragma unroll
      for (int I = 0; I < NTC; I++) {
        ps.fC[I] = panel_solve<p>(ch.Ct[I][J], ps.wb, false);
        const int R = 8 * I + r;
        if (R < b && col < b) { DBG_IDX(R * b + col, BB, "Z"); ch.Z[R * b + col] = ps.fC[I].e0; }
      }
    }
#pragma unroll
    for (int Jt = (h ? J + 1 : J); Jt < NTC; Jt++) {
      const double bD = (HAVE_X && Jt == X) ? ps.bDX : bfrag(ps.fD[Jt].e0, Jt);
#pragma unroll
      for (int I = Jt; I < NT; I++)
        if (!(HAVE_X && I == X && Jt == X)) dmma(ch.Dt[I][Jt], ps.fD[I].e1, bD);
#pragma unroll
      for (int I = 0; I <= J && I < NTC; I++) dmma(ch.Et[I][Jt], ps.fE[I].e1, bD);
      if (HASC) {
#pragma unroll
        for (int I = 0; I < NTC; I++) dmma(ch.Ct[I][Jt], ps.fC[I].e1, bD);
      }
    }
    if (HASC) {
      double bC[NTC];
#pragma unroll
      for (int Jn = 0; Jn < NTC; Jn++) bC[Jn] = __shfl_sync(FULL, ps.fC[Jn].e0, srcB);
#pragma unroll
      for (int I = 0; I < NT; I++) {
        double A = (I < NTC) ? ps.fC[I < NTC ? I : 0].e1 : 0.0;
        if (I == IR && r == RR) A = ps.fD[IR].e1;
#pragma unroll
        for (int Jn = 0; Jn <= I && Jn < NTC; Jn++) dmma(ch.Nt[I][Jn], A, bC[Jn]);
      }
    }
  }

  template <bool HASC, int p>
  __device__ __forceinline__ void la_step(Chain& ch, PanelState& prev, double* scr) const {
    __syncwarp();                                      // the gathered block of panel p is visible
    potrf4<p>(scr);
    if constexpr (p > 0) bulk<HASC, p - 1>(ch, prev);  // overlaps the factorization's latency chain
    __syncwarp();                                      // W is visible
    PanelState ps;
    crit<p>(ch, ps, scr);
    if constexpr (p + 1 < NP) la_step<HASC, p + 1>(ch, ps, scr);
    else bulk<HASC, p>(ch, ps);
  }
  template <bool HASC>
  __device__ __forceinline__ void block_la(Chain& ch, double* scr) const {
#pragma unroll
    for (int I = 0; I < NTC; I++)
#pragma unroll
      for (int J = 0; J < NTC; J++) { ch.Et[I][J].e0 = (I == J && r == j) ? 1.0 : 0.0; ch.Et[I][J].e1 = (I == J && r == j + 4) ? 1.0 : 0.0; }
    gather<0>(ch, scr);
    PanelState none;
    la_step<HASC, 0>(ch, none, scr);
  }

  // ---- the same solve by a TWO-WARP block, one elimination chain per warp (warp 0: blocks 0..m-1 downwards, the
  //      middle block and the back substitution of its half; warp 1: blocks N-1..m+1 upwards and its half).  Same
  //      shared-memory footprint per trajectory, twice the resident warps: the chains are long dependent sequences
  //      (26 k cycles per solve for a lone warp running both), so the SM needs warps, not instructions.
  //      Called by all 64 threads; contains __syncthreads().  scr: 2 * SCR doubles (SCR per warp).  Each warp only
  //      needs ITS half of H (and g) to be complete on entry; pre_middle() is run by warp 0 after the chains, before
  //      the middle block is loaded. ----
  template <class PreMiddle>
  __device__ __forceinline__ void solve2(double* Hd, double* Ho, const double* g, double* dl, double lambda, int N, double* scr, PreMiddle&& pre_middle) const {
    const int w = (threadIdx.x >> 5) & 1;
    const int m = N / 2, nd = m, nu = N - 1 - m;
    double* myscr = scr + w * SCR;
    myscr[lane] = 0.0;
    Chain ch[1];
    zero_N(ch[0]);
    const double lam0 = (r == j) ? lambda : 0.0, lam1 = (r == j + 4) ? lambda : 0.0;
    __syncwarp();
    const int nmine = w ? nu : nd;
    long long tq = clock64();
#pragma unroll 1
    for (int s = 0; s < nmine; s++) {
      const int i = w ? N - 1 - s : s;
      const int io = w ? i - 1 : i;                    // the coupling block towards the next block of the sweep
      load_D(ch[0], Hd + i * BD, g + i * b, lam0, lam1);
      load_C(ch[0], Ho + io * BB, w ? b : 1, w ? 1 : b);
      ch[0].L = Hd + i * BD; ch[0].Z = Ho + io * BB; ch[0].y = dl + i * b;
      __syncwarp();
      block_la<true>(ch[0], myscr);
    }
    MMA_PT(0, tq); tq = clock64();
    // warp 1 hands its Schur complement of the middle block over by adding it to H_mm in place, and the middle block's
    // rhs (g_m + its rhs update) through the y slot of the middle block (g itself stays intact for the caller)
    if (w == 1) {
      double* Hm = Hd + m * BD;
#pragma unroll
      for (int I = 0; I < NT; I++)
#pragma unroll
        for (int J = 0; J <= I && J < NTC; J++) {
          const int R = 8 * I + r;
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = 8 * J + 4 * e + j;
            const double v = e ? ch[0].Nt[I][J].e1 : ch[0].Nt[I][J].e0;      // (zero when this chain is empty)
            if (c < b && R < b && c <= R) { DBG_IDX(R * (R + 1) / 2 + c, BD, "H_mm"); Hm[R * (R + 1) / 2 + c] += v; }
            if (c < b && R == b) dl[m * b + c] = g[m * b + c] + v;
          }
        }
    }
    __syncthreads();
    MMA_PT(1, tq); tq = clock64();
    if (w == 0) {
      pre_middle();                                      // (the caller's last additions to H_mm and to the rhs in dl_m)
      load_D(ch[0], Hd + m * BD, dl + m * b, lam0, lam1);
      ch[0].L = Hd + m * BD; ch[0].Z = nullptr; ch[0].y = dl + m * b;
      __syncwarp();
      block_la<false>(ch[0], myscr);
      __syncwarp();
    }
    // ---- back substitution: warp 0 the middle block (s = 0), then each warp its chain outwards (one call site:
    //      the kernel is instruction-cache sensitive) ----
#pragma unroll 1
    for (int s = 0; s <= nmine; s++) {
      if (s > 0 || w == 0) back_block(Hd, Ho, dl, s == 0 ? m : (w ? m + s : m - s), s == 0 ? -1 : (w ? m + s - 1 : m - s), w ? m + s - 1 : m - s + 1);
      if (s == 0) __syncthreads();                       // x_m is in dl
    }
  }

  // one block of the back substitution:  x_i = L_ii^-T (y_i - Z^T x_prev), delta_i = -x_i, as two mat-vecs -- lane c forms
  // t_c = y_c + sum_R Z[R][c] delta_prev[R], the t go through the block's slot of dl, lane r forms x_r = sum_{c >= r}
  // U[r][c] t_c with the explicit U = L_ii^-T of the forward pass (no sequential triangular solve: a block is ~300 cycles
  // instead of ~900).  io: index of the coupling block holding Z (-1: the middle block, no coupling); ip: the block solved before.
  __device__ __forceinline__ void back_block(const double* Hd, const double* Ho, double* dl, int i, int io, int ip) const {
    const int c = lane, cl = min(c, b - 1);
    const bool valid = c < b;
    double t = dl[i * b + cl];                           // y_i (forward-substituted +g)
    if (io >= 0) {
      const double* Zp = Ho + io * BB + cl;
      const double* xp = dl + ip * b;
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int R = 0; R < b; R += 2) {
        const double2 x2 = *reinterpret_cast<const double2*>(xp + R);
        t0 = fma(Zp[R * b], x2.x, t0);
        t1 = fma(Zp[(R + 1) * b], x2.y, t1);
      }
      t += t0 + t1;
      __syncwarp();                                      // (every lane has read its y before the slot is reused for t)
      if (valid) dl[i * b + c] = t;
      __syncwarp();
    }
    const double* Up = Hd + i * BD + cl;                 // U[r][c] at c (c + 1) / 2 + r: lane r walks the columns c >= r
    const double* tp = dl + i * b;
    double x0 = 0.0, x1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < b; cc += 2) {
      const double2 t2 = *reinterpret_cast<const double2*>(tp + cc);
      const double u0 = Up[min(cc * (cc + 1) / 2, BD - 1 - cl)], u1 = Up[min((cc + 1) * (cc + 2) / 2, BD - 1 - cl)];
      x0 = fma(cc >= c ? u0 : 0.0, t2.x, x0);
      x1 = fma(cc + 1 >= c ? u1 : 0.0, t2.y, x1);
    }
    __syncwarp();
    if (valid) { DBG_IDX(i * b + c, dbg_nb, "delta"); dl[i * b + c] = -(x0 + x1); }
    __syncwarp();
  }
};

}  // namespace mma
