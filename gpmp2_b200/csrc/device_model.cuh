// device_model.cuh -- per-configuration math, one configuration per LANE, everything in registers:
//   DH forward kinematics            (replaces gpmp2/kinematics/Arm.cpp:31-143, pose part)
//   mobile-base lift + arm chain     (replaces gpmp2/kinematics/Pose2MobileArm.cpp:30-108,
//                                     gpmp2/kinematics/mobileBaseUtils.cpp:18-48)
//   sphere centres + Jacobian rows   (replaces gpmp2/kinematics/RobotModel-inl.h:12-40)
//   SDF trilinear / bilinear lookup  (replaces gpmp2/obstacle/SignedDistanceField.h:93-167,
//                                     gpmp2/obstacle/PlanarSDF.h:59-116)
//   hinge obstacle cost              (replaces gpmp2/obstacle/ObstacleCost.h:26-78)
//   unary / GP obstacle factor error + whitened J^T J, J^T e of the factor
//                                    (replaces gpmp2/obstacle/ObstacleSDFFactor-inl.h:18-55,
//                                     ObstaclePlanarSDFFactor-inl.h:18-56 and the per-configuration part of
//                                     ObstacleSDFFactorGP-inl.h:18-75 / ObstaclePlanarSDFFactorGP-inl.h:18-78)
//
// Formulation (not the reference's): every joint is a line in Pluecker coordinates (z_k, m_k = o_k x z_k);
// for a sphere at p with SDF gradient f the hinge Jacobian entry is
//      d e / d q_k = -grad d . (z_k x (p - o_k)) = -( z_k . (p x f) + m_k . f ),
// which equals the reference's [-R c^, R] * (inv(T_link) dT_link/dq_k)^vee chain (checked against the
// oracle, which computes it the reference's way).  Prismatic pseudo-joints of the mobile base are lines
// with z_k = 0, m_k = direction.
#pragma once
#include <math_constants.h>
#include "kparams.h"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Compact sin/cos: Cody-Waite reduction by pi/2 (two fma terms, exact enough for |x| <= 1e5) and the
// fdlibm kernel polynomials on [-pi/4, pi/4]; ~1 ulp.  The library sincos (Payne-Hanek slow path, ~270
// SASS instructions per call site) is kept out of line for huge arguments only -- 21 inlined copies of it
// were the largest single item of the kernel's instruction-cache footprint.
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void sincos_huge(double x, double* s, double* c) { sincos(x, s, c); }
// polynomial coefficients in the constant bank: a DFMA can take c[][] as an operand, a 64-bit immediate
// costs two extra moves per use
static __constant__ double SC_S[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double SC_C[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                               -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
static __constant__ double SC_R[3] = {0.63661977236758134308, 1.57079632679489655800e+00, 6.12323399573676603587e-17};

__device__ __forceinline__ void fast_sincos(double x, double& sn, double& cs) {
  if (fabs(x) > 1.0e5) { sincos_huge(x, &sn, &cs); return; }
  const double kd = rint(x * SC_R[0]);
  const int k = __double2int_rn(kd);
  double r = fma(-kd, SC_R[1], x);
  r = fma(-kd, SC_R[2], r);
  const double z = r * r;
  double ps = SC_S[5];
  ps = fma(ps, z, SC_S[4]);
  ps = fma(ps, z, SC_S[3]);
  ps = fma(ps, z, SC_S[2]);
  ps = fma(ps, z, SC_S[1]);
  ps = fma(ps, z, SC_S[0]);
  const double sr = fma(ps * z, r, r);
  double pc = SC_C[5];
  pc = fma(pc, z, SC_C[4]);
  pc = fma(pc, z, SC_C[3]);
  pc = fma(pc, z, SC_C[2]);
  pc = fma(pc, z, SC_C[1]);
  pc = fma(pc, z, SC_C[0]);
  const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
  const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
  sn = (k & 2) ? -s0 : s0;
  cs = ((k + 1) & 2) ? -c0 : c0;
}

// ---------------------------------------------------------------------------------------------
// SDF lookups.  Value and gradient share ONE set of corner reads (the reference reads them twice).
// Out-of-range -> false (the reference throws SDFQueryOutOfRange, caught by the hinge as zero cost).
// Index = (z*cols + col)*rows + row, the reference's per-slice column-major storage.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sdf_fix_axis(int& l, double& f, int n) {
  if ((unsigned)l < (unsigned)(n - 1)) return true;
  if (l == n - 1 && f == 0.0) { l = n - 2; f = 1.0; return true; }
  return false;
}

#ifndef GPMP2B_SDF_NOALLOC
#define GPMP2B_SDF_NOALLOC 1
#endif
// one 256-bit read-only load (LDG.E.ENL2.256.CONSTANT): the four values of a quad cell
struct Quad { double v00, v10, v01, v11; };   // (row, col), (row+1, col), (row, col+1), (row+1, col+1)
__device__ __forceinline__ Quad ldg_quad(const double* p) {
  Quad q;
#if GPMP2B_SDF_NOALLOC
  asm("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(q.v00), "=d"(q.v10), "=d"(q.v01), "=d"(q.v11) : "l"(p));
#else
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(q.v00), "=d"(q.v10), "=d"(q.v01), "=d"(q.v11) : "l"(p));
#endif
  return q;
}

template <bool GRAD>
__device__ __forceinline__ bool sdf3_lookup(const KSdf& f, double px, double py, double pz, double& dist,
                                            double& gx, double& gy, double& gz) {
  const double col = (px - f.ox) * f.inv_cell;
  const double row = (py - f.oy) * f.inv_cell;
  const double zz = (pz - f.oz) * f.inv_cell;
  int lc = __double2int_rd(col), lr = __double2int_rd(row), lz = __double2int_rd(zz);
  double fc = col - (double)lc, fr = row - (double)lr, fz = zz - (double)lz;
  // Range test in cell coordinates (SignedDistanceField.h:105-110): floor in [0, n-2] is the common case, one
  // unsigned compare per axis.  A point exactly on the upper boundary has floor = n-1 and weight 0 on the
  // (non-existent) upper neighbour (the reference reads one past the end there, :129-131): use cell n-2 with
  // fraction 1, the same interpolated value.  Everything else is out of range.
  if (!(((unsigned)lc < (unsigned)(f.cols - 1)) & ((unsigned)lr < (unsigned)(f.rows - 1)) & ((unsigned)lz < (unsigned)(f.nz - 1)))) {
    if (!(sdf_fix_axis(lc, fc, f.cols) && sdf_fix_axis(lr, fr, f.rows) && sdf_fix_axis(lz, fz, f.nz))) return false;
  }
  // 32-bit cell index (the host refuses fields with >= 2^31 cells); two quad cells, slices lz and lz + 1
  const int R = f.rows, RC = f.rows * f.cols;
  DBG_IDX(lr, f.rows, "sdf row"); DBG_IDX(lc, f.cols, "sdf col"); DBG_IDX(lz + 1, f.nz, "sdf slice");
  const double* __restrict__ p0 = f.quad + 4 * (size_t)(unsigned)(lz * RC + lc * R + lr);
  const Quad q0 = ldg_quad(p0), q1 = ldg_quad(p0 + 4 * (size_t)(unsigned)RC);
  // v[r][c][z]
  const double v000 = q0.v00, v100 = q0.v10, v010 = q0.v01, v110 = q0.v11;
  const double v001 = q1.v00, v101 = q1.v10, v011 = q1.v01, v111 = q1.v11;
  // along row
  const double d00 = v100 - v000, d10 = v110 - v010, d01 = v101 - v001, d11 = v111 - v011;
  const double a00 = fma(fr, d00, v000), a10 = fma(fr, d10, v010), a01 = fma(fr, d01, v001), a11 = fma(fr, d11, v011);
  // along col
  const double e0 = a10 - a00, e1 = a11 - a01;
  const double b0 = fma(fc, e0, a00), b1 = fma(fc, e1, a01);
  // along z
  const double gzz = b1 - b0;
  dist = fma(fz, gzz, b0);
  if (GRAD) {
    const double gcc = fma(fz, e1 - e0, e0);
    const double r0 = fma(fc, d10 - d00, d00), r1 = fma(fc, d11 - d01, d01);
    const double grr = fma(fz, r1 - r0, r0);
    // metric gradient (x,y,z) = (g_col, g_row, g_z) / cell   (SignedDistanceField.h:97)
    gx = gcc * f.inv_cell;
    gy = grr * f.inv_cell;
    gz = gzz * f.inv_cell;
  }
  return true;
}

template <bool GRAD>
__device__ __forceinline__ bool sdf2_lookup(const KSdf& f, double px, double py, double& dist, double& gx, double& gy) {
  const double col = (px - f.ox) * f.inv_cell;
  const double row = (py - f.oy) * f.inv_cell;
  int lc = __double2int_rd(col), lr = __double2int_rd(row);
  double fc = col - (double)lc, fr = row - (double)lr;
  if (!(((unsigned)lc < (unsigned)(f.cols - 1)) & ((unsigned)lr < (unsigned)(f.rows - 1)))) {
    if (!(sdf_fix_axis(lc, fc, f.cols) && sdf_fix_axis(lr, fr, f.rows))) return false;
  }
  DBG_IDX(lr, f.rows, "sdf row"); DBG_IDX(lc, f.cols, "sdf col");
  const Quad q = ldg_quad(f.quad + 4 * (size_t)(unsigned)(lc * f.rows + lr));
  const double v00 = q.v00, v10 = q.v10, v01 = q.v01, v11 = q.v11;
  const double d0 = v10 - v00, d1 = v11 - v01;
  const double a0 = fma(fr, d0, v00), a1 = fma(fr, d1, v01);
  const double e = a1 - a0;
  dist = fma(fc, e, a0);
  if (GRAD) {
    gx = e * f.inv_cell;                          // d/dcol
    gy = fma(fc, d1 - d0, d0) * f.inv_cell;       // d/drow
  }
  return true;
}

// frame [X Y Z | o] <- frame * T, T = 3x4 row-major [R | t] (Pose3::compose)
__device__ __forceinline__ void frame_compose(double (&X)[3], double (&Y)[3], double (&Z)[3], double (&o)[3], const double* T) {
  double nX[3], nY[3], nZ[3], no[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    nX[k] = X[k] * T[0] + Y[k] * T[4] + Z[k] * T[8];
    nY[k] = X[k] * T[1] + Y[k] * T[5] + Z[k] * T[9];
    nZ[k] = X[k] * T[2] + Y[k] * T[6] + Z[k] * T[10];
    no[k] = fma(Z[k], T[11], fma(Y[k], T[7], fma(X[k], T[3], o[k])));
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { X[k] = nX[k]; Y[k] = nY[k]; Z[k] = nZ[k]; o[k] = no[k]; }
}

// ---------------------------------------------------------------------------------------------
// One collision-checked configuration, evaluated by ONE lane.  qf(d) returns tangent coordinate d of the
// configuration (support state or GP-interpolated state).
//   JAC : accumulate whitened M = sum_s r_s r_s^T (packed lower, T = D(D+1)/2), cv = sum_s r_s e_s
//   always: err2 = sum_s (e_s * inv_sigma)^2 ; esum = sum_s e_s (unwhitened, for CollisionCost)
//   DBG : dump unwhitened e_s and sphere centres (original sphere order)
// KIND 0: q = joint angles.  KIND 1: q = (x, y, theta, joints...), link 0 = vehicle base.
// The joint loop is deliberately NOT unrolled (one copy of the FK step and of the sphere body): the fully
// unrolled version was ~37 KB of SASS and the kernel was instruction-fetch bound (ncu: stall_no_inst 54%).
// The joint lines live in registers; line k is captured with predicated moves when the loop reaches joint k,
// and every loop over k is guarded by the warp-uniform count nj of joints the current link depends on.
// ---------------------------------------------------------------------------------------------
// MASKED (linearize kernel of the phase pipeline): `smask` bit s = sphere s (kernel order) can be inside its hinge at
// this configuration, as found by the error evaluation of the same states (config_error<MASK>); `wmask` = the OR of
// smask over the lanes of the warp.  A sphere outside every lane's mask is skipped by the whole warp (no centre, no
// SDF gather), the chain stops after the last link any lane needs, and a lane skips the spheres outside its own mask.
// The error pass marks with a margin (MASK_MARGIN), so every sphere whose hinge decision could go either way is still
// evaluated here and decided by this function's own arithmetic: M and cv are bit-identical to the unmasked call.
#define GPMP2B_MASK_MARGIN 1e-9
// GEN (KIND 1): the general Pose2Vector chain -- torso link behind a linear actuator, second arm, run-time number of base
// coordinates (Pose2Mobile2Arms, Pose2MobileVetLinArm, Pose2MobileVetLin2Arms); false: the plain Pose2MobileArm chain with
// everything about it known at compile time (the run-time fields cost config 4 about 4 %).
template <int D, int NDIM, int KIND, bool JAC, bool DBG, bool MASKED = false, bool GEN = false, class QF>
__device__ __forceinline__ void config_eval(const KRobot& rb, const KSdf& sdf, const QF& qf, double eps,
                                            double inv_sigma, double (&M)[D * (D + 1) / 2], double (&cv)[D],
                                            double& err2, double& esum, double* dbg_err, double* dbg_ctr,
                                            unsigned long long smask = ~0ull, unsigned long long wmask = ~0ull) {
  constexpr int NB = (KIND == 1) ? 3 : 0;   // pseudo-joints of the mobile base (Pose2 part)
  const int nb = (KIND == 1) ? (GEN ? rb.nb : 3) : 0;   // ... 4 with a linear actuator (Pose2MobileVetLin*: coordinate 3 = torso lift)
  int link0 = (KIND == 1) ? 1 : 0;          // link of the first DH joint (2 behind a torso link)
  double zax[D][3], mom[D][3];              // joint lines (JAC only; dead code otherwise)
  double X[3], Y[3], Z[3], o[3];
  int s = 0;
  if (JAC) {
#pragma unroll
    for (int k = 0; k < D; k++)
#pragma unroll
      for (int c = 0; c < 3; c++) { zax[k][c] = 0.0; mom[k][c] = 0.0; }
  }

  // ---- one sphere on the current link frame [X Y Z | o]; nj = number of joints it depends on ----
  auto sphere = [&](int nj) {
    if (MASKED && !((smask >> s) & 1ull)) return;
    const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = fma(Z[k], cz, fma(Y[k], cy, fma(X[k], cx, o[k])));
    const double total_eps = rb.sph_r[s] + eps;
    double dist, f[3] = {0.0, 0.0, 0.0};
    bool in;
    if (NDIM == 3) in = sdf3_lookup<JAC>(sdf, p[0], p[1], p[2], dist, f[0], f[1], f[2]);
    else in = sdf2_lookup<JAC>(sdf, p[0], p[1], dist, f[0], f[1]);
    const bool active = in && !(dist > total_eps);   // ObstacleCost.h:40: dist > eps -> zero
    const double e = active ? total_eps - dist : 0.0;
    if (DBG) {
      const int so = rb.sph_orig[s];
      dbg_err[so] = e;
      if (dbg_ctr) { dbg_ctr[3 * so] = p[0]; dbg_ctr[3 * so + 1] = p[1]; dbg_ctr[3 * so + 2] = p[2]; }
    }
    if (active) {
      const double ew = e * inv_sigma;
      err2 = fma(ew, ew, err2);
      esum += e;
      if (JAC) {
        // torque of the gradient about the world origin
        const double tx = p[1] * f[2] - p[2] * f[1];
        const double ty = p[2] * f[0] - p[0] * f[2];
        const double tz = p[0] * f[1] - p[1] * f[0];
        double row[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
          row[k] = 0.0;
          if (k < nj) {
            double r = mom[k][0] * f[0];
            r = fma(mom[k][1], f[1], r);
            r = fma(mom[k][2], f[2], r);
            if (!(KIND == 1 && k < 2)) {   // prismatic base directions have z = 0
              r = fma(zax[k][0], tx, r);
              r = fma(zax[k][1], ty, r);
              r = fma(zax[k][2], tz, r);
            }
            row[k] = -r * inv_sigma;
          }
        }
#pragma unroll
        for (int a = 0; a < D; a++) {
          if (a < nj) {
            cv[a] = fma(row[a], ew, cv[a]);
#pragma unroll
            for (int b = 0; b <= a; b++) M[a * (a + 1) / 2 + b] = fma(row[a], row[b], M[a * (a + 1) / 2 + b]);
          }
        }
      }
    }
  };

  // ---- base ----
  if (KIND == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      X[k] = rb.base[k * 4 + 0]; Y[k] = rb.base[k * 4 + 1]; Z[k] = rb.base[k * 4 + 2]; o[k] = rb.base[k * 4 + 3];
    }
  } else {
    // computeBasePose3: Rz(theta), t = (x, y, 0)
    double sn, cs;
    fast_sincos(qf(2), sn, cs);
    X[0] = cs; X[1] = sn; X[2] = 0.0;
    Y[0] = -sn; Y[1] = cs; Y[2] = 0.0;
    Z[0] = 0.0; Z[1] = 0.0; Z[2] = 1.0;
    o[0] = qf(0); o[1] = qf(1); o[2] = 0.0;
    if (JAC) {
#pragma unroll
      for (int k = 0; k < 3; k++) { zax[0][k] = 0.0; mom[0][k] = X[k]; zax[1][k] = 0.0; mom[1][k] = Y[k]; }
      zax[2][0] = 0.0; zax[2][1] = 0.0; zax[2][2] = 1.0;
      mom[2][0] = o[1]; mom[2][1] = -o[0]; mom[2][2] = 0.0;    // o x z
    }
    // spheres on the vehicle (link 0)
    for (const int se = rb.sph_begin[1]; s < se; s++) {
      if (MASKED && !((wmask >> s) & 1ull)) continue;
      sphere(3);
    }
    // arm base = vehicle * base_T_arm (computeBaseTransPose3) ...
    frame_compose(X, Y, Z, o, rb.base);
    if (GEN && rb.lift) {
      // ... or the torso link of Pose2MobileVetLinArm / VetLin2Arms: Trans(0, 0, +-z) * vehicle * base_T_torso
      // (liftBasePose3, mobileBaseUtils.cpp:51-82); its coordinate is a translation along the world z axis
      o[2] = fma((double)rb.lift, qf(3), o[2]);
      if (JAC) {
#pragma unroll
        for (int k = NB; k < D; k++)
          if (k == 3) { zax[k][0] = 0.0; zax[k][1] = 0.0; zax[k][2] = 0.0; mom[k][0] = 0.0; mom[k][1] = 0.0; mom[k][2] = (double)rb.lift; }
      }
      for (const int se = rb.sph_begin[2]; s < se; s++) sphere(4);
      frame_compose(X, Y, Z, o, rb.base2);   // arm 1 base = torso * torso_T_arm1
      link0 = 2;
    }
  }

  // ---- DH chain: T_{j+1} = T_j Rz(q_j + bias_j) Trans(a_j, 0, d_j) Rx(alpha_j)  (Arm.cpp:24-27, Arm.h:93-98)
  int jend = D - nb;
  if (MASKED && KIND == 0) {   // the last link any lane of the warp needs (wmask != 0: the caller skips the call otherwise)
    const int link_hi = rb.sph_link[63 - __clzll((long long)wmask)];
    jend = min(jend, link_hi + 1);
  }
#pragma unroll 1
  for (int j = 0; j < jend; j++) {
    if (GEN && KIND == 1 && j == rb.n1) {
      // second arm (Pose2Mobile2Arms.cpp:78-101, Pose2MobileVetLin2Arms.cpp:88-111): its chain starts again from the
      // vehicle (or torso) frame, and its spheres do not move with arm 1's joints -- their lines are cleared, so the
      // rows of those coordinates come out as zeros without touching the sphere body
      double sn, cs;
      fast_sincos(qf(2), sn, cs);
      X[0] = cs; X[1] = sn; X[2] = 0.0;
      Y[0] = -sn; Y[1] = cs; Y[2] = 0.0;
      Z[0] = 0.0; Z[1] = 0.0; Z[2] = 1.0;
      o[0] = qf(0); o[1] = qf(1); o[2] = 0.0;
      if (rb.lift) {
        frame_compose(X, Y, Z, o, rb.base);
        o[2] = fma((double)rb.lift, qf(3), o[2]);
        frame_compose(X, Y, Z, o, rb.base3);
      } else {
        frame_compose(X, Y, Z, o, rb.base2);
      }
      if (JAC) {
#pragma unroll
        for (int k = NB; k < D; k++)
          if (k >= nb && k < nb + j) {
#pragma unroll
            for (int c = 0; c < 3; c++) { zax[k][c] = 0.0; mom[k][c] = 0.0; }
          }
      }
    }
    if (JAC) {
      const double m0 = o[1] * Z[2] - o[2] * Z[1], m1 = o[2] * Z[0] - o[0] * Z[2], m2 = o[0] * Z[1] - o[1] * Z[0];
#pragma unroll
      for (int k = NB; k < D; k++)
        if (k == nb + j) {
          zax[k][0] = Z[0]; zax[k][1] = Z[1]; zax[k][2] = Z[2];
          mom[k][0] = m0; mom[k][1] = m1; mom[k][2] = m2;
        }
    }
    double sn, cs;
    fast_sincos(qf(nb + j) + rb.bias[j], sn, cs);
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
    const int link = link0 + j;
#pragma unroll 1
    for (const int se = rb.sph_begin[link + 1]; s < se; s++) {
      if (MASKED && !((wmask >> s) & 1ull)) continue;
      sphere(nb + j + 1);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// config_eval<JAC> with ASYNCHRONOUS SDF gathers (linearize kernel of the phase pipeline; arms in 3-D fields).
// In config_eval a lane waits for the gather of every sphere (one L2 / DRAM round trip per sphere: 33 % of the
// linearize kernel's samples sat on the first use of the corner values).  Here the configuration is walked TWICE:
// an issue walk (frame only) runs `chunk` spheres ahead, computes centres and cell addresses and issues the corner
// reads as 16-byte cp.async copies into the lane's own shared-memory slots (no registers held); the Jacobian walk
// (frame + joint lines) consumes a chunk once its copies have landed, while the next chunk's are in flight (two
// buffers).  Same arithmetic, in the same order, as config_eval + sdf3_lookup<true>: bit-identical M, cv, errors.
// scratch: 2 * chunk * 384 doubles per warp; slot layout as in config_error (parts 0..3 = the two quads,
// 4 = (fr, fc), 5 = (fz, eps')).
// ---------------------------------------------------------------------------------------------
template <int NDIM>
__device__ __forceinline__ bool sdf_cell(const KSdf& f, double px, double py, double pz, unsigned& cell, double& fr,
                                         double& fc, double& fz);

template <int D, class QF>
__device__ __forceinline__ void config_eval_async(const KRobot& rb, const KSdf& sdf, const QF& qf, double eps, double inv_sigma,
                                                  double (&M)[D * (D + 1) / 2], double (&cv)[D], double& err2, double& esum,
                                                  double* scratch, int chunk) {
  double zax[D][3], mom[D][3];
  double X[3], Y[3], Z[3], o[3];          // Jacobian walk
  double X1[3], Y1[3], Z1[3], o1[3];      // issue walk
#pragma unroll
  for (int k = 0; k < D; k++)
#pragma unroll
    for (int c = 0; c < 3; c++) { zax[k][c] = 0.0; mom[k][c] = 0.0; }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    X[k] = X1[k] = rb.base[k * 4 + 0]; Y[k] = Y1[k] = rb.base[k * 4 + 1]; Z[k] = Z1[k] = rb.base[k * 4 + 2]; o[k] = o1[k] = rb.base[k * 4 + 3];
  }
  int link1 = -1, link2 = -1;
  const int S = rb.n_spheres;
  double2* sl = reinterpret_cast<double2*>(scratch) + (threadIdx.x & 31);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sl);
  const size_t zstride = (size_t)(unsigned)(sdf.rows * sdf.cols) * 32;   // bytes between slices z and z + 1
  const int nch = (S + chunk - 1) / chunk;

  auto dh_step = [&](int j, double (&A)[3], double (&B)[3], double (&C)[3], double (&t)[3]) {
    double sn, cs;
    fast_sincos(qf(j) + rb.bias[j], sn, cs);
    const double ca = rb.ca[j], sa = rb.sa[j], aj = rb.a[j], dj = rb.d[j];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double xn = fma(cs, A[k], sn * B[k]);
      const double yn = fma(cs, B[k], -sn * A[k]);
      t[k] = fma(dj, C[k], fma(aj, xn, t[k]));
      const double y2 = fma(ca, yn, sa * C[k]);
      const double z2 = fma(ca, C[k], -sa * yn);
      A[k] = xn; B[k] = y2; C[k] = z2;
    }
  };
  auto issue = [&](int c) {
    const int s0 = c * chunk, ns = min(chunk, S - s0);
    const unsigned bufb = sbase + (unsigned)((c & 1) * chunk * 6 * 32 * 16);
    double2* slb = sl + (c & 1) * chunk * 6 * 32;
#pragma unroll 1
    for (int u = 0; u < ns; u++) {
      const int s = s0 + u;
      const int link = rb.sph_link[s];
#pragma unroll 1
      while (link1 < link) { link1++; dh_step(link1, X1, Y1, Z1, o1); }
      const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
      double p[3];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = fma(Z1[k], cz, fma(Y1[k], cy, fma(X1[k], cx, o1[k])));
      unsigned cell;
      double fr, fc, fz;
      const bool in = sdf_cell<3>(sdf, p[0], p[1], p[2], cell, fr, fc, fz);
      const char* src = reinterpret_cast<const char*>(sdf.quad) + (size_t)cell * 32;
      const unsigned dst = bufb + (unsigned)(u * 6 * 32 * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u), "l"(src + 16) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 1024u), "l"(src + zstride) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 1536u), "l"(src + zstride + 16) : "memory");
      slb[(u * 6 + 4) * 32] = make_double2(fr, fc);
      slb[(u * 6 + 5) * 32] = make_double2(fz, in ? rb.sph_r[s] + eps : -CUDART_INF);   // out of range: never active
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(0);
#pragma unroll 1
  for (int c = 0; c < nch; c++) {
    if (c + 1 < nch) {
      issue(c + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const int s0 = c * chunk, ns = min(chunk, S - s0);
    const double2* qb = sl + (c & 1) * chunk * 6 * 32;
#pragma unroll 1
    for (int u = 0; u < ns; u++) {
      const int s = s0 + u;
      const int link = rb.sph_link[s];
#pragma unroll 1
      while (link2 < link) {
        link2++;
        const double m0 = o[1] * Z[2] - o[2] * Z[1], m1 = o[2] * Z[0] - o[0] * Z[2], m2 = o[0] * Z[1] - o[1] * Z[0];
#pragma unroll
        for (int k = 0; k < D; k++)
          if (k == link2) {
            zax[k][0] = Z[0]; zax[k][1] = Z[1]; zax[k][2] = Z[2];
            mom[k][0] = m0; mom[k][1] = m1; mom[k][2] = m2;
          }
        dh_step(link2, X, Y, Z, o);
      }
      const int nj = link + 1;
      const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
      double p[3];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = fma(Z[k], cz, fma(Y[k], cy, fma(X[k], cx, o[k])));
      const double2* q = qb + u * 6 * 32;
      const double2 a0 = q[0], a1 = q[32], b0 = q[2 * 32], b1 = q[3 * 32], w0 = q[4 * 32], w1 = q[5 * 32];
      const double fr = w0.x, fc = w0.y, fz = w1.x, total_eps = w1.y;
      // value and gradient: the arithmetic of sdf3_lookup<true>
      const double v000 = a0.x, v100 = a0.y, v010 = a1.x, v110 = a1.y;
      const double v001 = b0.x, v101 = b0.y, v011 = b1.x, v111 = b1.y;
      const double d00 = v100 - v000, d10 = v110 - v010, d01 = v101 - v001, d11 = v111 - v011;
      const double a00 = fma(fr, d00, v000), a10 = fma(fr, d10, v010), a01 = fma(fr, d01, v001), a11 = fma(fr, d11, v011);
      const double e0 = a10 - a00, e1 = a11 - a01;
      const double c0 = fma(fc, e0, a00), c1 = fma(fc, e1, a01);
      const double gzz = c1 - c0;
      const double dist = fma(fz, gzz, c0);
      const double gcc = fma(fz, e1 - e0, e0);
      const double r0 = fma(fc, d10 - d00, d00), r1 = fma(fc, d11 - d01, d01);
      const double grr = fma(fz, r1 - r0, r0);
      double f[3];
      f[0] = gcc * sdf.inv_cell; f[1] = grr * sdf.inv_cell; f[2] = gzz * sdf.inv_cell;
      const bool active = !(dist > total_eps);   // ObstacleCost.h:40 (eps' = -inf for an out-of-range sphere)
      if (active) {
        const double e = total_eps - dist;
        const double ew = e * inv_sigma;
        err2 = fma(ew, ew, err2);
        esum += e;
        const double tx = p[1] * f[2] - p[2] * f[1];
        const double ty = p[2] * f[0] - p[0] * f[2];
        const double tz = p[0] * f[1] - p[1] * f[0];
        double row[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
          row[k] = 0.0;
          if (k < nj) {
            double r = mom[k][0] * f[0];
            r = fma(mom[k][1], f[1], r);
            r = fma(mom[k][2], f[2], r);
            r = fma(zax[k][0], tx, r);
            r = fma(zax[k][1], ty, r);
            r = fma(zax[k][2], tz, r);
            row[k] = -r * inv_sigma;
          }
        }
#pragma unroll
        for (int a = 0; a < D; a++) {
          if (a < nj) {
            cv[a] = fma(row[a], ew, cv[a]);
#pragma unroll
            for (int bb = 0; bb <= a; bb++) M[a * (a + 1) / 2 + bb] = fma(row[a], row[bb], M[a * (a + 1) / 2 + bb]);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Error-only evaluation of one configuration: FK + sphere centres + SDF VALUE + hinge, no Jacobians.  Used by the
// candidate-error pass of LM (1.7 evaluations per iteration) and CollisionCost.  A lone warp spent ~1000 cycles per
// sphere waiting for its L2 gather when spheres were looked up one at a time; the lookup is split into ISSUE (cell
// address + fractions; out-of-range lanes read cell 0) and FINISH (interpolation).  Two ways to keep gathers in flight:
// asynchronous copies into a shared-memory scratch lent by the caller (the fast path), or one sphere of software
// pipelining in registers.  More registers in flight (2, 4, 8 spheres) always lost: the loop body must stay small.
// ---------------------------------------------------------------------------------------------
#ifndef GPMP2B_ERR_SMEM
#define GPMP2B_ERR_SMEM 1
#endif
#ifndef GPMP2B_ERR_QD      // register path of the error pass: spheres whose gathers are in flight per lane
#define GPMP2B_ERR_QD 1
#endif
template <int NDIM>
struct SdfTap {
  double v[NDIM == 3 ? 8 : 4];
  double fr, fc, fz;
  bool in;
  bool in_list = false;   // pipelined error pass: a real sphere (not an empty pipeline slot)
};

// cell index (clamped to cell 0 when the point is outside) and interpolation fractions of a lookup
template <int NDIM>
__device__ __forceinline__ bool sdf_cell(const KSdf& f, double px, double py, double pz, unsigned& cell, double& fr,
                                         double& fc, double& fz) {
  const double col = (px - f.ox) * f.inv_cell;
  const double row = (py - f.oy) * f.inv_cell;
  int lc = __double2int_rd(col), lr = __double2int_rd(row), lz = 0;
  fc = col - (double)lc;
  fr = row - (double)lr;
  fz = 0.0;
  bool ok = ((unsigned)lc < (unsigned)(f.cols - 1)) & ((unsigned)lr < (unsigned)(f.rows - 1));
  if (NDIM == 3) {
    const double zz = (pz - f.oz) * f.inv_cell;
    lz = __double2int_rd(zz);
    fz = zz - (double)lz;
    ok &= (unsigned)lz < (unsigned)(f.nz - 1);
  }
  if (!ok) {   // rare: on the upper boundary, or outside
    ok = sdf_fix_axis(lc, fc, f.cols) && sdf_fix_axis(lr, fr, f.rows);
    if (NDIM == 3) ok = ok && sdf_fix_axis(lz, fz, f.nz);
    if (!ok) { lc = 0; lr = 0; lz = 0; }
  }
  DBG_IDX(lr, f.rows, "sdf row"); DBG_IDX(lc, f.cols, "sdf col");
  if (NDIM == 3) DBG_IDX(lz + 1, f.nz, "sdf slice");
  cell = (unsigned)((NDIM == 3 ? lz * (f.rows * f.cols) : 0) + lc * f.rows + lr);
  return ok;
}

template <int NDIM>
__device__ __forceinline__ void sdf_issue(const KSdf& f, double px, double py, double pz, SdfTap<NDIM>& t) {
  unsigned cell;
  t.in = sdf_cell<NDIM>(f, px, py, pz, cell, t.fr, t.fc, t.fz);
  const double* __restrict__ p0 = f.quad + 4 * (size_t)cell;
  if (NDIM == 3) {
    const Quad q0 = ldg_quad(p0), q1 = ldg_quad(p0 + 4 * (size_t)(unsigned)(f.rows * f.cols));
    t.v[0] = q0.v00; t.v[1] = q0.v10; t.v[2] = q0.v01; t.v[3] = q0.v11;
    t.v[4] = q1.v00; t.v[5] = q1.v10; t.v[6] = q1.v01; t.v[7] = q1.v11;
  } else {
    const Quad q = ldg_quad(p0);
    t.v[0] = q.v00; t.v[1] = q.v10; t.v[2] = q.v01; t.v[3] = q.v11;
  }
}

template <int NDIM>
__device__ __forceinline__ double sdf_finish_value(const SdfTap<NDIM>& t) {
  const double fr = t.fr, fc = t.fc;
  if (NDIM == 3) {
    const double a00 = fma(fr, t.v[1] - t.v[0], t.v[0]), a10 = fma(fr, t.v[3] - t.v[2], t.v[2]);
    const double a01 = fma(fr, t.v[5] - t.v[4], t.v[4]), a11 = fma(fr, t.v[7] - t.v[6], t.v[6]);
    const double b0 = fma(fc, a10 - a00, a00), b1 = fma(fc, a11 - a01, a01);
    return fma(t.fz, b1 - b0, b0);
  } else {
    const double a0 = fma(fr, t.v[1] - t.v[0], t.v[0]), a1 = fma(fr, t.v[3] - t.v[2], t.v[2]);
    return fma(fc, a1 - a0, a0);
  }
}

// MASK: *amask collects bit s for every sphere within MASK_MARGIN of its hinge (see config_eval<MASKED>); register path only.
template <int D, int NDIM, int KIND, bool DBG, bool MASK = false, bool GEN = false, class QF>
__device__ __forceinline__ void config_error(const KRobot& rb, const KSdf& sdf, const QF& qf, double eps, double inv_sigma,
                                             double& err2, double& esum, double* dbg_err, double* dbg_ctr,
                                             double* scratch = nullptr, int chunk = 0, unsigned long long* amask = nullptr) {
  const int nb = (KIND == 1) ? (GEN ? rb.nb : 3) : 0;
  double X[3], Y[3], Z[3], o[3];
  int link_cur;
  auto vehicle = [&]() {   // computeBasePose3: Rz(theta), t = (x, y, 0)
    double sn, cs;
    fast_sincos(qf(2), sn, cs);
    X[0] = cs; X[1] = sn; X[2] = 0.0;
    Y[0] = -sn; Y[1] = cs; Y[2] = 0.0;
    Z[0] = 0.0; Z[1] = 0.0; Z[2] = 1.0;
    o[0] = qf(0); o[1] = qf(1); o[2] = 0.0;
  };
  if (KIND == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      X[k] = rb.base[k * 4 + 0]; Y[k] = rb.base[k * 4 + 1]; Z[k] = rb.base[k * 4 + 2]; o[k] = rb.base[k * 4 + 3];
    }
    link_cur = -1;
  } else {
    vehicle();
    link_cur = 0;
  }
  const int S = rb.n_spheres;
  const int link0 = (KIND == 1) ? ((GEN && rb.lift) ? 2 : 1) : 0;   // link of the first DH joint
  // advance the kinematic chain to `link` (one DH step per iteration; Arm.cpp:24-27, Pose2MobileArm.cpp:30-108; the
  // torso link and the second arm of Pose2Mobile2Arms.cpp / Pose2MobileVetLinArm.cpp / Pose2MobileVetLin2Arms.cpp)
  auto advance = [&](int link) {
#pragma unroll 1
    while (link_cur < link) {
      if (KIND == 1) {
        if (link_cur == 0) {                     // arm base = vehicle * base_T_arm, or the torso before its lift
          frame_compose(X, Y, Z, o, rb.base);
          if (GEN && rb.lift) { o[2] = fma((double)rb.lift, qf(3), o[2]); link_cur = 1; continue; }   // torso link: no DH step
        } else if (GEN && rb.lift && link_cur == 1) {   // arm 1 base = torso * torso_T_arm1
          frame_compose(X, Y, Z, o, rb.base2);
        } else if (GEN && link_cur == link0 + rb.n1 - 1 && rb.n1 < rb.arm_dof) {   // last link of arm 1 -> arm 2 starts from its own base
          vehicle();
          if (rb.lift) {
            frame_compose(X, Y, Z, o, rb.base);
            o[2] = fma((double)rb.lift, qf(3), o[2]);
            frame_compose(X, Y, Z, o, rb.base3);
          } else {
            frame_compose(X, Y, Z, o, rb.base2);
          }
        }
      }
      link_cur++;
      const int j = link_cur - link0;
      double sn, cs;
      fast_sincos(qf(nb + j) + rb.bias[j], sn, cs);
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
  };
#if GPMP2B_ERR_SMEM
  if (!DBG && !MASK && scratch != nullptr) {
    // Asynchronous gathers through shared memory.  The caller lends the (dead) H storage as scratch: for a chunk of
    // spheres every lane issues its quad-cell reads as 16-byte cp.async copies straight into its own slots (all of
    // them in flight together, no registers held), keeps the interpolation fractions next to them, waits ONCE, and
    // interpolates.  One L2 round trip per chunk instead of one per sphere; both loops are rolled (one copy of code).
    // Slot layout (double2 units): [(u * 6 + part) * 32 + lane], parts 0..3 = the two quads, 4 = (fr, fc), 5 = (fz, eps').
    double2* sl = reinterpret_cast<double2*>(scratch) + (threadIdx.x & 31);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sl);
    const size_t zstride = (size_t)(unsigned)(sdf.rows * sdf.cols) * 32;   // bytes between slices z and z + 1
#pragma unroll 1
    for (int s0 = 0; s0 < S; s0 += chunk) {
      const int ns = min(chunk, S - s0);
#pragma unroll 1
      for (int u = 0; u < ns; u++) {
        const int s = s0 + u;
        advance(rb.sph_link[s]);
        const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
        double p[3];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = fma(Z[k], cz, fma(Y[k], cy, fma(X[k], cx, o[k])));
        unsigned cell;
        double fr, fc, fz;
        const bool in = sdf_cell<NDIM>(sdf, p[0], p[1], p[2], cell, fr, fc, fz);
        const char* src = reinterpret_cast<const char*>(sdf.quad) + (size_t)cell * 32;
        const unsigned dst = sbase + (unsigned)(u * 6 * 32 * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u), "l"(src + 16) : "memory");
        if (NDIM == 3) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 1024u), "l"(src + zstride) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 1536u), "l"(src + zstride + 16) : "memory");
        }
        // an out-of-range sphere gets eps' = -inf: "dist > eps'" is then always true -> zero cost (ObstacleCost.h:33-38)
        sl[(u * 6 + 4) * 32] = make_double2(fr, fc);
        sl[(u * 6 + 5) * 32] = make_double2(fz, in ? rb.sph_r[s] + eps : -CUDART_INF);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll 1
      for (int u = 0; u < ns; u++) {
        const double2* q = sl + u * 6 * 32;
        SdfTap<NDIM> t;
        const double2 a0 = q[0], a1 = q[32], m0 = q[4 * 32], m1 = q[5 * 32];
        t.v[0] = a0.x; t.v[1] = a0.y; t.v[2] = a1.x; t.v[3] = a1.y;
        if (NDIM == 3) {
          const double2 b0 = q[2 * 32], b1 = q[3 * 32];
          t.v[4] = b0.x; t.v[5] = b0.y; t.v[6] = b1.x; t.v[7] = b1.y;
        }
        t.fr = m0.x; t.fc = m0.y; t.fz = m1.x;
        const double dist = sdf_finish_value<NDIM>(t);
        const double e = !(dist > m1.y) ? m1.y - dist : 0.0;   // ObstacleCost.h:40
        const double ew = e * inv_sigma;
        err2 = fma(ew, ew, err2);
        esum += e;
      }
    }
    return;
  }
#endif
  // Register path (debug entry, systems too small for the scratch).  Software pipeline over the spheres, ONE copy of the chain-advance / centre / issue / finish code (the hot
  // code must stay small): the gather of sphere s is consumed QD iterations later, so its L2
  // round trip overlaps the forward kinematics and address arithmetic of the following spheres.
  constexpr int QD = GPMP2B_ERR_QD;
  SdfTap<NDIM> pend[QD];
  double peps[QD];
  int pidx[QD];
#pragma unroll
  for (int u = 0; u < QD; u++) {
#pragma unroll
    for (int k = 0; k < (NDIM == 3 ? 8 : 4); k++) pend[u].v[k] = 0.0;
    pend[u].fr = pend[u].fc = pend[u].fz = 0.0;
    pend[u].in = false;
    peps[u] = 0.0;
    pidx[u] = 0;
  }
  auto finish = [&](const SdfTap<NDIM>& t, double te, int si) {
    const double dist = sdf_finish_value<NDIM>(t);
    const bool active = t.in && !(dist > te);   // ObstacleCost.h:40
    const double e = active ? te - dist : 0.0;
    const double ew = e * inv_sigma;
    err2 = fma(ew, ew, err2);
    esum += e;
    if (MASK) { if (t.in && !(dist > te + GPMP2B_MASK_MARGIN)) *amask |= 1ull << si; }
    if (DBG && t.in_list) dbg_err[rb.sph_orig[si]] = e;
  };
#pragma unroll 1
  for (int s = 0; s < S; s++) {
    const int link = rb.sph_link[s];
    advance(link);
    const double cx = rb.sph_c[s][0], cy = rb.sph_c[s][1], cz = rb.sph_c[s][2];
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = fma(Z[k], cz, fma(Y[k], cy, fma(X[k], cx, o[k])));
    SdfTap<NDIM> nw;
    sdf_issue<NDIM>(sdf, p[0], p[1], p[2], nw);
    nw.in_list = true;
    const double neps = rb.sph_r[s] + eps;
    if (DBG && dbg_ctr) {
      const int so = rb.sph_orig[s];
      dbg_ctr[3 * so] = p[0]; dbg_ctr[3 * so + 1] = p[1]; dbg_ctr[3 * so + 2] = p[2];
    }
    finish(pend[0], peps[0], pidx[0]);
#pragma unroll
    for (int u = 0; u + 1 < QD; u++) { pend[u] = pend[u + 1]; peps[u] = peps[u + 1]; pidx[u] = pidx[u + 1]; }
    pend[QD - 1] = nw; peps[QD - 1] = neps; pidx[QD - 1] = s;
  }
#pragma unroll
  for (int u = 0; u < QD; u++) finish(pend[u], peps[u], pidx[u]);
}
