// pose2.cuh -- the SE(2) formulas of GTSAM's Pose2 that gpmp2 reaches through traits<Pose2>:
// between / compose / Logmap / Expmap, their derivatives and AdjointMap.  Used by the Pose2Vector optimizer
// (optimizer_kernel_lie.cuh) and by the trajectory utilities (c_abi.cu: straight-line init, GP densification).
#pragma once
#include "device_model.cuh"

namespace p2 {
struct Pose { double x, y, th; };
__device__ __forceinline__ double wrap_pi(double t) { return t - 6.283185307179586477 * rint(t * 0.15915494309189533577); }
// a^-1 b   (Pose2::between)
__device__ __forceinline__ Pose between(const Pose& a, const Pose& b) {
  double s, c;
  fast_sincos(a.th, s, c);
  const double dx = b.x - a.x, dy = b.y - a.y;
  return Pose{fma(c, dx, s * dy), fma(c, dy, -s * dx), wrap_pi(b.th - a.th)};
}
// a b      (Pose2::compose)
__device__ __forceinline__ Pose compose(const Pose& a, const Pose& b) {
  double s, c;
  fast_sincos(a.th, s, c);
  return Pose{a.x + fma(c, b.x, -s * b.y), a.y + fma(s, b.x, c * b.y), wrap_pi(a.th + b.th)};
}
// Pose2::Logmap
__device__ __forceinline__ void logmap(const Pose& p, double (&v)[3]) {
  const double w = p.th;
  if (fabs(w) < 1e-10) { v[0] = p.x; v[1] = p.y; v[2] = w; return; }
  double s, c;
  fast_sincos(w, s, c);
  const double c_1 = c - 1.0, det = c_1 * c_1 + s * s;
  const double ux = fma(c, p.x, s * p.y), uy = fma(c, p.y, -s * p.x);   // unrotate(t)
  const double dx = ux - p.x, dy = uy - p.y;
  v[0] = (w / det) * (-dy);
  v[1] = (w / det) * dx;
  v[2] = w;
}
// Pose2::LogmapDerivative (row-major 3x3)
__device__ __forceinline__ void logmap_derivative(const double (&v)[3], double (&J)[9]) {
  const double alpha = v[2];
#pragma unroll
  for (int k = 0; k < 9; k++) J[k] = 0.0;
  if (fabs(alpha) > 1e-5) {
    double s, c;
    fast_sincos(alpha, s, c);
    const double alphaInv = 1.0 / alpha, h = 0.5 * s / (1.0 - c);
    J[0] = alpha * h; J[1] = -0.5 * alpha; J[2] = v[0] * alphaInv - v[0] * h + 0.5 * v[1];
    J[3] = 0.5 * alpha; J[4] = alpha * h;  J[5] = v[1] * alphaInv - 0.5 * v[0] - v[1] * h;
    J[8] = 1.0;
  } else {
    J[0] = 1.0; J[2] = 0.5 * v[1];
    J[4] = 1.0; J[5] = -0.5 * v[0];
    J[8] = 1.0;
  }
}
// Pose2::Expmap
__device__ __forceinline__ Pose expmap(const double (&v)[3]) {
  const double w = v[2];
  if (fabs(w) < 1e-10) return Pose{v[0], v[1], w};
  double s, c;
  fast_sincos(w, s, c);
  const double ox = -v[1], oy = v[0];
  return Pose{(ox - (c * ox - s * oy)) / w, (oy - (s * ox + c * oy)) / w, wrap_pi(w)};
}
// Pose2::ExpmapDerivative
__device__ __forceinline__ void expmap_derivative(const double (&v)[3], double (&J)[9]) {
  const double alpha = v[2];
#pragma unroll
  for (int k = 0; k < 9; k++) J[k] = 0.0;
  if (fabs(alpha) > 1e-5) {
    double s, c;
    fast_sincos(alpha, s, c);
    const double sZ = s / alpha, cZ = (c - 1.0) / alpha, v1Z = v[0] / alpha, v2Z = v[1] / alpha;
    J[0] = sZ; J[1] = -cZ; J[2] = v1Z + v2Z * cZ - v1Z * sZ;
    J[3] = cZ; J[4] = sZ;  J[5] = -v1Z * cZ + v2Z - v2Z * sZ;
    J[8] = 1.0;
  } else {
    J[0] = 1.0; J[2] = -0.5 * v[1];
    J[4] = 1.0; J[5] = 0.5 * v[0];
    J[8] = 1.0;
  }
}
// Pose2::AdjointMap of p^-1
__device__ __forceinline__ void adjoint_of_inverse(const Pose& p, double (&A)[9]) {
  double s, c;
  fast_sincos(p.th, s, c);
  // p^-1 = (-(c x + s y), -(-s x + c y), -th); Ad(q) = [[cq, -sq, qy], [sq, cq, -qx], [0, 0, 1]]
  const double ix = -fma(c, p.x, s * p.y), iy = -fma(c, p.y, -s * p.x);
  A[0] = c; A[1] = s; A[2] = iy;
  A[3] = -s; A[4] = c; A[5] = -ix;
  A[6] = 0.0; A[7] = 0.0; A[8] = 1.0;
}
__device__ __forceinline__ void mul33(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[i * 3 + j] = fma(A[i * 3 + 2], B[6 + j], fma(A[i * 3 + 1], B[3 + j], A[i * 3] * B[j]));
}
}  // namespace p2
