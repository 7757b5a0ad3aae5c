// gpmp2_oracle.cpp -- CPU restatement (fp64) of the GPMP2 batched trajectory-optimization path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpmp2_b200/, include/) links, imports or
// calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do, as the checker / the timed CPU baseline.
//
// Every function cites the reference file:line (relative to the ori-drs/gpmp2 tree) it restates.
// The factor arithmetic (kinematics, SDF, hinge, GP) is gpmp2 code and is pinned by the reference's
// own unit-test vectors (tests/golden/, tests/test_oracle_golden.py).  The optimizer underneath
// (whitening, normal equations, LM lambda logic, retract, checkConvergence) is GTSAM 4.0.x
// (branch `wrap-export`, reference README.md:13) which is NOT vendored in the reference tree and
// cannot be built here: that part is restated from GTSAM's published algorithm
// (LevenbergMarquardtOptimizer::iterate/tryLambda, NonlinearOptimizer checkConvergence) and is
// "PARITY UNPINNED" at the end-to-end level -- the reference has no planner test or fixture
// (last row).  Only the 2-state Gauss-Newton checks of the reference tests pin it.
//
// Deliberately written the way the reference computes (4x4 homogeneous FK, se(3) pose Jacobians,
// explicit per-factor whitened Jacobian blocks, explicit Lambda/Psi matrices) so that the CUDA
// path, which uses different but algebraically equal formulations, is checked independently.

#include "../include/gpmp2b.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <thread>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------
// tiny dense matrix (row-major)
// ------------------------------------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  static Mat Identity(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; i++) m(i, i) = 1.0;
    return m;
  }
};
typedef std::vector<double> Vec;

static Mat mul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; i++)
    for (int k = 0; k < A.c; k++) {
      const double aik = A(i, k);
      if (aik == 0.0) continue;
      for (int j = 0; j < B.c; j++) C(i, j) += aik * B(k, j);
    }
  return C;
}
static Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; i++)
    for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
  return T;
}
static Mat add(const Mat& A, const Mat& B, double sb = 1.0) {
  Mat C(A.r, A.c);
  for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] + sb * B.a[i];
  return C;
}
static Mat scale(const Mat& A, double s) {
  Mat C(A.r, A.c);
  for (size_t i = 0; i < A.a.size(); i++) C.a[i] = s * A.a[i];
  return C;
}
static Vec mulv(const Mat& A, const Vec& x) {
  Vec y(A.r, 0.0);
  for (int i = 0; i < A.r; i++) {
    double s = 0;
    for (int j = 0; j < A.c; j++) s += A(i, j) * x[j];
    y[i] = s;
  }
  return y;
}
static Mat block(const Mat& A, int r0, int c0, int nr, int nc) {
  Mat B(nr, nc);
  for (int i = 0; i < nr; i++)
    for (int j = 0; j < nc; j++) B(i, j) = A(r0 + i, c0 + j);
  return B;
}
static void set_block(Mat& A, int r0, int c0, const Mat& B) {
  for (int i = 0; i < B.r; i++)
    for (int j = 0; j < B.c; j++) A(r0 + i, c0 + j) = B(i, j);
}
// general inverse, Gauss-Jordan with partial pivoting (Eigen's dynamic inverse() is PartialPivLU)
static Mat inverse(const Mat& A) {
  const int n = A.r;
  Mat M = A, I = Mat::Identity(n);
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(M(i, k)) > std::fabs(M(p, k))) p = i;
    if (M(p, k) == 0.0) throw std::runtime_error("[oracle] singular matrix");
    if (p != k)
      for (int j = 0; j < n; j++) {
        std::swap(M(k, j), M(p, j));
        std::swap(I(k, j), I(p, j));
      }
    const double d = 1.0 / M(k, k);
    for (int j = 0; j < n; j++) {
      M(k, j) *= d;
      I(k, j) *= d;
    }
    for (int i = 0; i < n; i++) {
      if (i == k) continue;
      const double f = M(i, k);
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) {
        M(i, j) -= f * M(k, j);
        I(i, j) -= f * I(k, j);
      }
    }
  }
  return I;
}
// upper Cholesky R with R^T R = A  (GTSAM RtR(), used by noiseModel::Gaussian::Information)
static Mat chol_upper(const Mat& A) {
  const int n = A.r;
  Mat R(n, n);
  for (int j = 0; j < n; j++) {
    double s = A(j, j);
    for (int k = 0; k < j; k++) s -= R(k, j) * R(k, j);
    if (s <= 0.0) throw std::runtime_error("[oracle] chol_upper: not SPD");
    R(j, j) = std::sqrt(s);
    for (int i = j + 1; i < n; i++) {
      double t = A(j, i);
      for (int k = 0; k < j; k++) t -= R(k, j) * R(k, i);
      R(j, i) = t / R(j, j);
    }
  }
  return R;
}

// ------------------------------------------------------------------------------------------------
// rigid transforms as 4x4 row-major (what gtsam::Pose3::matrix() returns)
// ------------------------------------------------------------------------------------------------
typedef std::array<double, 16> M4;
static M4 m4_identity() {
  M4 m{};
  m[0] = m[5] = m[10] = m[15] = 1.0;
  return m;
}
static M4 m4_mul(const M4& A, const M4& B) {
  M4 C{};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = s;
    }
  return C;
}
static M4 m4_rigid_inverse(const M4& A) {  // Pose3::inverse(): (R^T, -R^T t)
  M4 B = m4_identity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i * 4 + j] = A[j * 4 + i];
  for (int i = 0; i < 3; i++)
    B[i * 4 + 3] = -(B[i * 4 + 0] * A[3] + B[i * 4 + 1] * A[7] + B[i * 4 + 2] * A[11]);
  return B;
}
static M4 m4_trans(double x, double y, double z) {
  M4 m = m4_identity();
  m[3] = x; m[7] = y; m[11] = z;
  return m;
}
static M4 m4_rx(double t) {  // Rot3::Rx
  M4 m = m4_identity();
  const double c = std::cos(t), s = std::sin(t);
  m[5] = c; m[6] = -s; m[9] = s; m[10] = c;
  return m;
}
static M4 m4_rz(double t) {  // Rot3::Rz
  M4 m = m4_identity();
  const double c = std::cos(t), s = std::sin(t);
  m[0] = c; m[1] = -s; m[4] = s; m[5] = c;
  return m;
}

// ------------------------------------------------------------------------------------------------
// SE(2) pieces used by the mobile-arm path [GTSAM Pose2, recalled -- SURVEY.md App. B.4]
// ------------------------------------------------------------------------------------------------
struct Pose2 {
  double x = 0, y = 0, th = 0;
};
static Pose2 p2_compose(const Pose2& a, const Pose2& b) {
  const double c = std::cos(a.th), s = std::sin(a.th);
  Pose2 r;
  r.x = a.x + c * b.x - s * b.y;
  r.y = a.y + s * b.x + c * b.y;
  r.th = a.th + b.th;
  // GTSAM keeps Rot2 as (c,s) and theta() = atan2(s,c): wrap to (-pi, pi]
  r.th = std::atan2(std::sin(r.th), std::cos(r.th));
  return r;
}
static Pose2 p2_inverse(const Pose2& a) {
  const double c = std::cos(a.th), s = std::sin(a.th);
  Pose2 r;
  r.x = -(c * a.x + s * a.y);
  r.y = -(-s * a.x + c * a.y);
  r.th = -a.th;
  return r;
}
static Mat p2_adjoint(const Pose2& p) {  // Pose2::AdjointMap
  const double c = std::cos(p.th), s = std::sin(p.th);
  Mat A(3, 3);
  A(0, 0) = c; A(0, 1) = -s; A(0, 2) = p.y;
  A(1, 0) = s; A(1, 1) = c;  A(1, 2) = -p.x;
  A(2, 2) = 1.0;
  return A;
}
static Pose2 p2_expmap(const double v[3]) {  // Pose2::Expmap
  const double w = v[2];
  Pose2 r;
  if (std::fabs(w) < 1e-10) {
    r.x = v[0]; r.y = v[1]; r.th = w;
  } else {
    const double c = std::cos(w), s = std::sin(w);
    // t = (v_ortho - R v_ortho)/w, v_ortho = (-vy, vx)
    const double ox = -v[1], oy = v[0];
    r.x = (ox - (c * ox - s * oy)) / w;
    r.y = (oy - (s * ox + c * oy)) / w;
    r.th = std::atan2(s, c);
  }
  return r;
}
static void p2_logmap(const Pose2& p, double v[3]) {  // Pose2::Logmap
  const double w = p.th;
  if (std::fabs(w) < 1e-10) {
    v[0] = p.x; v[1] = p.y; v[2] = w;
  } else {
    const double c_1 = std::cos(w) - 1.0, s = std::sin(w);
    const double det = c_1 * c_1 + s * s;
    // p = R_PI_2 * (R^T t - t) ; R_PI_2 (a,b) = (-b, a)
    const double c = std::cos(w);
    const double ux = c * p.x + s * p.y, uy = -s * p.x + c * p.y;  // unrotate(t)
    const double dx = ux - p.x, dy = uy - p.y;
    const double px = -dy, py = dx;
    v[0] = (w / det) * px;
    v[1] = (w / det) * py;
    v[2] = w;
  }
}
static Mat p2_expmap_derivative(const double v[3]) {  // Pose2::ExpmapDerivative
  const double alpha = v[2];
  Mat J(3, 3);
  if (std::fabs(alpha) > 1e-5) {
    const double sZalpha = std::sin(alpha) / alpha, c_1Zalpha = (std::cos(alpha) - 1) / alpha;
    const double v1Zalpha = v[0] / alpha, v2Zalpha = v[1] / alpha;
    J(0, 0) = sZalpha;   J(0, 1) = -c_1Zalpha; J(0, 2) = v1Zalpha + v2Zalpha * c_1Zalpha - v1Zalpha * sZalpha;
    J(1, 0) = c_1Zalpha; J(1, 1) = sZalpha;    J(1, 2) = -v1Zalpha * c_1Zalpha + v2Zalpha - v2Zalpha * sZalpha;
    J(2, 2) = 1.0;
  } else {
    J(0, 0) = 1; J(0, 2) = -0.5 * v[1];
    J(1, 1) = 1; J(1, 2) = 0.5 * v[0];
    J(2, 2) = 1;
  }
  return J;
}
static Mat p2_logmap_derivative(const Pose2& p) {  // Pose2::LogmapDerivative
  double v[3];
  p2_logmap(p, v);
  const double alpha = v[2];
  Mat J(3, 3);
  if (std::fabs(alpha) > 1e-5) {
    const double alphaInv = 1 / alpha;
    const double halfCotHalfAlpha = 0.5 * std::sin(alpha) / (1 - std::cos(alpha));
    const double v1 = v[0], v2 = v[1];
    J(0, 0) = alpha * halfCotHalfAlpha; J(0, 1) = -0.5 * alpha;
    J(0, 2) = v1 * alphaInv - v1 * halfCotHalfAlpha + 0.5 * v2;
    J(1, 0) = 0.5 * alpha; J(1, 1) = alpha * halfCotHalfAlpha;
    J(1, 2) = v2 * alphaInv - 0.5 * v1 - v2 * halfCotHalfAlpha;
    J(2, 2) = 1;
  } else {
    J(0, 0) = 1; J(0, 2) = 0.5 * v[1];
    J(1, 1) = 1; J(1, 2) = -0.5 * v[0];
    J(2, 2) = 1;
  }
  return J;
}

// ------------------------------------------------------------------------------------------------
// robot model
// ------------------------------------------------------------------------------------------------
struct Robot {
  int kind = GPMP2B_ROBOT_ARM;
  int arm_dof = 0;
  int dof = 0;       // system dof (arm: arm_dof; mobile: 3 + [1 lift] + arm_dof + arm2_dof)
  int nr_links = 0;  // arm: arm_dof; mobile: 1 + [1 torso] + arm_dof + arm2_dof
  int arm2_dof = 0;  // second arm (Pose2Mobile2Arms, Pose2MobileVetLin2Arms)
  bool lift = false, reverse_linact = false;   // linear actuator (Pose2MobileVetLinArm, Pose2MobileVetLin2Arms)
  Vec a, alpha, d, bias;   // arm 1 joints, then arm 2 joints
  M4 base;                       // ARM: base pose; MOBILE: base_T_arm / base_T_arm1 / base_T_torso
  M4 base2, base3;               // base_T_arm2 | torso_T_arm | torso_T_arm1 ; torso_T_arm2  (include/gpmp2b.h)
  std::vector<M4> link_notheta;  // gpmp2/kinematics/Arm.cpp:16-28
  std::vector<int> sph_link;
  Vec sph_radius, sph_center;

  explicit Robot(const gpmp2b_robot_desc& r) {
    kind = r.kind;
    arm_dof = r.arm_dof;
    arm2_dof = (kind == GPMP2B_ROBOT_POSE2_MOBILE_2ARMS || kind == GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_2ARMS) ? r.arm2_dof : 0;
    lift = kind == GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_ARM || kind == GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_2ARMS;
    reverse_linact = lift && r.reverse_linact != 0;
    const int nj = arm_dof + arm2_dof;
    dof = kind == GPMP2B_ROBOT_ARM ? arm_dof : nj + 3 + (lift ? 1 : 0);
    nr_links = kind == GPMP2B_ROBOT_ARM ? arm_dof : nj + 1 + (lift ? 1 : 0);
    a.assign(r.a, r.a + nj);
    alpha.assign(r.alpha, r.alpha + nj);
    d.assign(r.d, r.d + nj);
    if (r.theta_bias) bias.assign(r.theta_bias, r.theta_bias + nj);
    else bias.assign(nj, 0.0);
    for (int i = 0; i < 16; i++) { base[i] = r.base_pose[i]; base2[i] = r.base_pose2[i]; base3[i] = r.base_pose3[i]; }
    // Arm::Arm, gpmp2/kinematics/Arm.cpp:16-28: Trans(0,0,d) * Trans(a,0,0) * Rx(alpha)
    for (int i = 0; i < nj; i++)
      link_notheta.push_back(m4_mul(m4_mul(m4_trans(0, 0, d[i]), m4_trans(a[i], 0, 0)), m4_rx(alpha[i])));
    sph_link.assign(r.sphere_link, r.sphere_link + r.n_spheres);
    sph_radius.assign(r.sphere_radius, r.sphere_radius + r.n_spheres);
    sph_center.assign(r.sphere_center, r.sphere_center + 3 * r.n_spheres);
  }
  int nr_spheres() const { return (int)sph_link.size(); }
};

// Arm::forwardKinematics, gpmp2/kinematics/Arm.cpp:31-143 (pose part; jv = none as the obstacle
// factors call it, RobotModel-inl.h:20-22).  J_pose[i] is 6 x arm_dof, rows [omega; v] body frame.
// The arm is the `dof` DH joints starting at joint `first` of the robot's tables (arm 1: first = 0; arm 2: first = arm_dof).
static void arm_fk(const Robot& rb, const M4& base_pose, const double* jp, std::vector<M4>& jpx,
                   std::vector<Mat>* J_jpx_jp, int first = 0, int dof = -1) {
  if (dof < 0) dof = rb.arm_dof;
  jpx.resize(dof);
  if (J_jpx_jp) J_jpx_jp->assign(dof, Mat(6, dof));
  std::vector<M4> H(dof), Ho(dof + 1), dH(dof), Hoinv(dof + 1);
  Ho[0] = base_pose;
  Hoinv[0] = m4_rigid_inverse(Ho[0]);
  for (int i = 1; i <= dof; i++) {
    // getH, Arm.h:93-98: Pose3(Rz(theta + bias), 0) * link_trans_notheta
    H[i - 1] = m4_mul(m4_rz(jp[i - 1] + rb.bias[first + i - 1]), rb.link_notheta[first + i - 1]);
    Ho[i] = m4_mul(Ho[i - 1], H[i - 1]);
    if (J_jpx_jp) {
      // getdH, Arm.h:101-110
      const double c = std::cos(jp[i - 1] + rb.bias[first + i - 1]), s = std::sin(jp[i - 1] + rb.bias[first + i - 1]);
      M4 dRot{};
      dRot[0] = -s; dRot[1] = -c; dRot[4] = c; dRot[5] = -s;
      dH[i - 1] = m4_mul(dRot, rb.link_notheta[first + i - 1]);
      Hoinv[i] = m4_rigid_inverse(Ho[i]);
    }
  }
  // dHo_dq cache, Arm.cpp:85-92
  std::vector<std::vector<M4>> dHo_dq(dof, std::vector<M4>(dof));
  if (J_jpx_jp) {
    for (int i = 0; i < dof; i++)
      for (int j = 0; j <= i; j++) {
        if (i > j) dHo_dq[i][j] = m4_mul(m4_mul(m4_mul(Ho[j], dH[j]), Hoinv[j + 1]), Ho[i + 1]);
        else dHo_dq[i][j] = m4_mul(Ho[j], dH[j]);
      }
  }
  for (int i = 0; i < dof; i++) {
    jpx[i] = Ho[i + 1];
    if (J_jpx_jp) {
      Mat& Jp = (*J_jpx_jp)[i];
      const M4 inv_jpx_i = m4_rigid_inverse(jpx[i]);
      for (int j = 0; j <= i; j++) {
        const M4 S = m4_mul(inv_jpx_i, dHo_dq[i][j]);  // Arm.cpp:110-113
        Jp(0, j) = S[2 * 4 + 1];
        Jp(1, j) = S[0 * 4 + 2];
        Jp(2, j) = S[1 * 4 + 0];
        Jp(3, j) = S[0 * 4 + 3];
        Jp(4, j) = S[1 * 4 + 3];
        Jp(5, j) = S[2 * 4 + 3];
      }
    }
  }
}

// Pose3::AdjointMap for tangent order [omega; v]: [[R,0],[skew(t) R, R]]  [GTSAM-recalled]
static Mat pose3_adjoint(const M4& T) {
  Mat A(6, 6);
  double R[3][3], t[3] = {T[3], T[7], T[11]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T[i * 4 + j];
  const double S[3][3] = {{0, -t[2], t[1]}, {t[2], 0, -t[0]}, {-t[1], t[0], 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A(i, j) = R[i][j];
      A(3 + i, 3 + j) = R[i][j];
      double s = 0;
      for (int k = 0; k < 3; k++) s += S[i][k] * R[k][j];
      A(3 + i, j) = s;
    }
  return A;
}

// Forward kinematics of the whole robot: link poses + optional 6 x dof pose Jacobians.
//  ARM    : Arm::forwardKinematics with the arm's own base pose.
//  MOBILE : Pose2MobileArm::forwardKinematics, gpmp2/kinematics/Pose2MobileArm.cpp:30-108, with
//           computeBasePose3 / computeBaseTransPose3, gpmp2/kinematics/mobileBaseUtils.cpp:18-48;
//           Pose2Mobile2Arms.cpp:33-104, Pose2MobileVetLinArm.cpp:31-94 and Pose2MobileVetLin2Arms.cpp:34-114 with
//           liftBasePose3 (mobileBaseUtils.cpp:51-82) are the same chain with a second arm and / or a lifted torso link.
static void robot_fk(const Robot& rb, const double* conf, std::vector<M4>& px, std::vector<Mat>* J) {
  if (rb.kind == GPMP2B_ROBOT_ARM) {
    arm_fk(rb, rb.base, conf, px, J);
    return;
  }
  const int n1 = rb.arm_dof, n2 = rb.arm2_dof, dof = rb.dof, nl = rb.nr_links;
  px.resize(nl);
  if (J) J->assign(nl, Mat(6, dof));
  // computeBasePose3: Pose3(Rodrigues(0,0,theta), (x,y,0)); J: rows 0-2 col 2 = ExpmapDerivative col 2
  // = (0,0,1); rows 3-4 cols 0-1 = I2   (mobileBaseUtils.cpp:18-31)
  M4 veh = m4_rz(conf[2]);
  veh[3] = conf[0]; veh[7] = conf[1]; veh[11] = 0.0;
  Mat Hveh(6, 3);
  Hveh(2, 2) = 1.0;
  Hveh(3, 0) = 1.0;
  Hveh(4, 1) = 1.0;
  px[0] = veh;
  if (J) set_block((*J)[0], 0, 0, Hveh);
  // computeBaseTransPose3: base_pose3.compose(base_T_trans, Hcomp); J = Hcomp * Hbasep3, Hcomp =
  // Ad(base_T_trans^-1)   (mobileBaseUtils.cpp:34-48)
  const int nb = rb.lift ? 4 : 3;      // columns of the base part: Pose2 (+ lift)
  const int L0 = rb.lift ? 2 : 1;      // first arm link
  M4 root = m4_mul(veh, rb.base);      // arm base (kinds 1, 2: arm 1) or torso before the lift
  Mat Hroot(6, nb);
  set_block(Hroot, 0, 0, mul(pose3_adjoint(m4_rigid_inverse(rb.base)), Hveh));
  if (rb.lift) {
    // liftBasePose3 (mobileBaseUtils.cpp:51-82): lift_base_pose.compose(armbase, Hcomp1, Hcomp2) with lift_base_pose =
    // Trans(0, 0, +-z); Hcomp2 = I, Hcomp1 = Ad(armbase^-1), d/dz = +-Hcomp1.col(5) = +-[0; R_armbase^T e_z]
    const double z = rb.reverse_linact ? -conf[3] : conf[3];
    root[11] += z;
    const double sg = rb.reverse_linact ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++) Hroot(3 + i, 3) = sg * root[2 * 4 + i];   // row 2 of R = column 2 of R^T... (R^T e_z)_i = R[2][i]
    px[1] = root;
    if (J) set_block((*J)[1], 0, 0, Hroot);
  }
  // bases of the arms and their Jacobians over the base part
  const bool two = n2 > 0;
  M4 arm_base[2];
  Mat Harm[2] = {Mat(6, nb), Mat(6, nb)};
  if (rb.lift) {
    // tso_base.compose(torso_T_arm, H_tso_comp), Harm_base = H_tso_comp * Htso_base  (Pose2MobileVetLinArm.cpp:58-62)
    arm_base[0] = m4_mul(root, rb.base2);
    Harm[0] = mul(pose3_adjoint(m4_rigid_inverse(rb.base2)), Hroot);
    if (two) { arm_base[1] = m4_mul(root, rb.base3); Harm[1] = mul(pose3_adjoint(m4_rigid_inverse(rb.base3)), Hroot); }
  } else {
    arm_base[0] = root;
    Harm[0] = Hroot;
    if (two) {   // computeBaseTransPose3(p.pose(), base_T_arm2_, Harm2_base)  (Pose2Mobile2Arms.cpp:58-60)
      arm_base[1] = m4_mul(veh, rb.base2);
      Harm[1] = mul(pose3_adjoint(m4_rigid_inverse(rb.base2)), Hveh);
    }
  }
  for (int arm = 0; arm < (two ? 2 : 1); arm++) {
    const int first = arm == 0 ? 0 : n1, n = arm == 0 ? n1 : n2;
    std::vector<M4> armjpx;
    std::vector<Mat> Jarm;
    arm_fk(rb, arm_base[arm], conf + nb + first, armjpx, J ? &Jarm : nullptr, first, n);
    for (int i = 0; i < n; i++) {
      px[L0 + first + i] = armjpx[i];
      if (J) {
        // Pose2MobileArm.cpp:100-101, Pose2Mobile2Arms.cpp:88-101, Pose2MobileVetLinArm.cpp:86-91
        const Mat Ad = pose3_adjoint(m4_mul(m4_rigid_inverse(armjpx[i]), arm_base[arm]));
        set_block((*J)[L0 + first + i], 0, 0, mul(Ad, Harm[arm]));
        set_block((*J)[L0 + first + i], 0, nb + first, Jarm[i]);
      }
    }
  }
}

// RobotModel<FK>::sphereCenters, gpmp2/kinematics/RobotModel-inl.h:12-40.
// centers: S x 3; J: S matrices 3 x dof.  Pose3::transform_from Jacobian = [-R*skew(c), R].
static void sphere_centers(const Robot& rb, const double* conf, Vec& centers, std::vector<Mat>* J) {
  std::vector<M4> poses;
  std::vector<Mat> Jpose;
  robot_fk(rb, conf, poses, J ? &Jpose : nullptr);
  const int S = rb.nr_spheres();
  centers.assign(3 * S, 0.0);
  if (J) J->assign(S, Mat(3, rb.dof));
  for (int s = 0; s < S; s++) {
    const M4& T = poses[rb.sph_link[s]];
    const double* c = &rb.sph_center[3 * s];
    for (int i = 0; i < 3; i++)
      centers[3 * s + i] = T[i * 4 + 0] * c[0] + T[i * 4 + 1] * c[1] + T[i * 4 + 2] * c[2] + T[i * 4 + 3];
    if (J) {
      Mat Jpp(3, 6);
      const double K[3][3] = {{0, -c[2], c[1]}, {c[2], 0, -c[0]}, {-c[1], c[0], 0}};
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          double sum = 0;
          for (int k = 0; k < 3; k++) sum += T[i * 4 + k] * K[k][j];
          Jpp(i, j) = -sum;
          Jpp(i, 3 + j) = T[i * 4 + j];
        }
      (*J)[s] = mul(Jpp, Jpose[rb.sph_link[s]]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// signed distance fields
// ------------------------------------------------------------------------------------------------
struct Sdf {
  int ndim, rows, cols, nz;
  double origin[3], cell;
  const double* data;
  explicit Sdf(const gpmp2b_sdf_desc& s)
      : ndim(s.ndim), rows(s.rows), cols(s.cols), nz(s.ndim == 3 ? s.nz : 1), cell(s.cell_size), data(s.data) {
    for (int i = 0; i < 3; i++) origin[i] = s.origin[i];
  }
  // data_[z](r, c), SignedDistanceField.h:170-172 (column-major Eigen matrix per slice)
  inline double at(long r, long c, long z) const {
    // the reference reads one past the end (weight 0) when a point sits exactly on the upper
    // boundary; clamp the index so the read is defined, the weight is 0.
    if (r >= rows) r = rows - 1;
    if (c >= cols) c = cols - 1;
    if (z >= nz) z = nz - 1;
    return data[((size_t)z * cols + c) * rows + r];
  }
};

// SignedDistanceField::getSignedDistance(point, g), gpmp2/obstacle/SignedDistanceField.h:93-99,
// with convertPoint3toCell :103-116, signed_distance :127-141, gradient :146-167.
// returns false when SDFQueryOutOfRange would be thrown.
static bool sdf3_query(const Sdf& f, const double p[3], double& dist, double g[3]) {
  if (p[0] < f.origin[0] || p[0] > (f.origin[0] + (f.cols - 1.0) * f.cell) ||
      p[1] < f.origin[1] || p[1] > (f.origin[1] + (f.rows - 1.0) * f.cell) ||
      p[2] < f.origin[2] || p[2] > (f.origin[2] + (f.nz - 1.0) * f.cell))
    return false;
  const double col = (p[0] - f.origin[0]) / f.cell;
  const double row = (p[1] - f.origin[1]) / f.cell;
  const double z = (p[2] - f.origin[2]) / f.cell;
  const double lr = std::floor(row), lc = std::floor(col), lz = std::floor(z);
  const double hr = lr + 1.0, hc = lc + 1.0, hz = lz + 1.0;
  const long lri = (long)lr, lci = (long)lc, lzi = (long)lz, hri = (long)hr, hci = (long)hc, hzi = (long)hz;
#define SD(r, c, zz) f.at(r, c, zz)
  const double gr =
      (hc - col) * (hz - z) * (SD(hri, lci, lzi) - SD(lri, lci, lzi)) +
      (col - lc) * (hz - z) * (SD(hri, hci, lzi) - SD(lri, hci, lzi)) +
      (hc - col) * (z - lz) * (SD(hri, lci, hzi) - SD(lri, lci, hzi)) +
      (col - lc) * (z - lz) * (SD(hri, hci, hzi) - SD(lri, hci, hzi));
  const double gc =
      (hr - row) * (hz - z) * (SD(lri, hci, lzi) - SD(lri, lci, lzi)) +
      (row - lr) * (hz - z) * (SD(hri, hci, lzi) - SD(hri, lci, lzi)) +
      (hr - row) * (z - lz) * (SD(lri, hci, hzi) - SD(lri, lci, hzi)) +
      (row - lr) * (z - lz) * (SD(hri, hci, hzi) - SD(hri, lci, hzi));
  const double gz =
      (hr - row) * (hc - col) * (SD(lri, lci, hzi) - SD(lri, lci, lzi)) +
      (row - lr) * (hc - col) * (SD(hri, lci, hzi) - SD(hri, lci, lzi)) +
      (hr - row) * (col - lc) * (SD(lri, hci, hzi) - SD(lri, hci, lzi)) +
      (row - lr) * (col - lc) * (SD(hri, hci, hzi) - SD(hri, hci, lzi));
  // SignedDistanceField.h:97: Vector3(g_idx(1), g_idx(0), g_idx(2)) / cell_size
  g[0] = gc / f.cell;
  g[1] = gr / f.cell;
  g[2] = gz / f.cell;
  dist = (hr - row) * (hc - col) * (hz - z) * SD(lri, lci, lzi) +
         (row - lr) * (hc - col) * (hz - z) * SD(hri, lci, lzi) +
         (hr - row) * (col - lc) * (hz - z) * SD(lri, hci, lzi) +
         (row - lr) * (col - lc) * (hz - z) * SD(hri, hci, lzi) +
         (hr - row) * (hc - col) * (z - lz) * SD(lri, lci, hzi) +
         (row - lr) * (hc - col) * (z - lz) * SD(hri, lci, hzi) +
         (hr - row) * (col - lc) * (z - lz) * SD(lri, hci, hzi) +
         (row - lr) * (col - lc) * (z - lz) * SD(hri, hci, hzi);
#undef SD
  return true;
}

// PlanarSDF::getSignedDistance(point, g), gpmp2/obstacle/PlanarSDF.h:59-65, :69-80, :90-116.
static bool sdf2_query(const Sdf& f, const double p[2], double& dist, double g[2]) {
  if (p[0] < f.origin[0] || p[0] > (f.origin[0] + (f.cols - 1.0) * f.cell) ||
      p[1] < f.origin[1] || p[1] > (f.origin[1] + (f.rows - 1.0) * f.cell))
    return false;
  const double col = (p[0] - f.origin[0]) / f.cell;
  const double row = (p[1] - f.origin[1]) / f.cell;
  const double lr = std::floor(row), lc = std::floor(col);
  const double hr = lr + 1.0, hc = lc + 1.0;
  const long lri = (long)lr, lci = (long)lc, hri = (long)hr, hci = (long)hc;
#define SD(r, c) f.at(r, c, 0)
  const double gr = (hc - col) * (SD(hri, lci) - SD(lri, lci)) + (col - lc) * (SD(hri, hci) - SD(lri, hci));
  const double gc = (hr - row) * (SD(lri, hci) - SD(lri, lci)) + (row - lr) * (SD(hri, hci) - SD(hri, lci));
  g[0] = gc / f.cell;
  g[1] = gr / f.cell;
  dist = (hr - row) * (hc - col) * SD(lri, lci) + (row - lr) * (hc - col) * SD(hri, lci) +
         (hr - row) * (col - lc) * SD(lri, hci) + (row - lr) * (col - lc) * SD(hri, hci);
#undef SD
  return true;
}

// hingeLossObstacleCost, gpmp2/obstacle/ObstacleCost.h:26-50 (3-D) and :54-78 (planar).
// H_point (1 x 3; planar uses the first two) may be NULL.
static double hinge_obstacle(const Sdf& f, const double p[3], double eps, double* H_point) {
  double dist, g[3] = {0, 0, 0};
  const bool ok = f.ndim == 3 ? sdf3_query(f, p, dist, g) : sdf2_query(f, p, dist, g);
  if (!ok || dist > eps) {
    if (H_point) H_point[0] = H_point[1] = H_point[2] = 0.0;
    return 0.0;
  }
  if (H_point) {
    H_point[0] = -g[0];
    H_point[1] = -g[1];
    H_point[2] = f.ndim == 3 ? -g[2] : 0.0;
  }
  return eps - dist;
}

// ObstacleSDFFactor<ROBOT>::evaluateError, gpmp2/obstacle/ObstacleSDFFactor-inl.h:18-55 and
// ObstaclePlanarSDFFactor<ROBOT>::evaluateError, ObstaclePlanarSDFFactor-inl.h:18-56 (the planar
// variant uses sphere (x,y) and J.topRows<2>(): the third component of H_point is 0 above).
static Vec obstacle_factor(const Robot& rb, const Sdf& f, const double* conf, double epsilon, Mat* H1) {
  const int S = rb.nr_spheres();
  Vec centers, err(S);
  std::vector<Mat> Jpx;
  sphere_centers(rb, conf, centers, H1 ? &Jpx : nullptr);
  if (H1) *H1 = Mat(S, rb.dof);
  for (int s = 0; s < S; s++) {
    const double total_eps = rb.sph_radius[s] + epsilon;
    if (H1) {
      double Jerr[3];
      err[s] = hinge_obstacle(f, &centers[3 * s], total_eps, Jerr);
      for (int j = 0; j < rb.dof; j++)
        (*H1)(s, j) = Jerr[0] * Jpx[s](0, j) + Jerr[1] * Jpx[s](1, j) + Jerr[2] * Jpx[s](2, j);
    } else {
      err[s] = hinge_obstacle(f, &centers[3 * s], total_eps, nullptr);
    }
  }
  return err;
}

// ------------------------------------------------------------------------------------------------
// GP utilities, gpmp2/gp/GPutils.h:25-59
// ------------------------------------------------------------------------------------------------
static Mat calcQ(const Mat& Qc, double tau) {
  const int d = Qc.r;
  Mat Q(2 * d, 2 * d);
  set_block(Q, 0, 0, scale(Qc, 1.0 / 3 * std::pow(tau, 3.0)));
  set_block(Q, 0, d, scale(Qc, 1.0 / 2 * std::pow(tau, 2.0)));
  set_block(Q, d, 0, scale(Qc, 1.0 / 2 * std::pow(tau, 2.0)));
  set_block(Q, d, d, scale(Qc, tau));
  return Q;
}
static Mat calcQ_inv(const Mat& Qc, double tau) {
  const int d = Qc.r;
  const Mat Qc_inv = inverse(Qc);
  Mat Q(2 * d, 2 * d);
  set_block(Q, 0, 0, scale(Qc_inv, 12.0 * std::pow(tau, -3.0)));
  set_block(Q, 0, d, scale(Qc_inv, (-6.0) * std::pow(tau, -2.0)));
  set_block(Q, d, 0, scale(Qc_inv, (-6.0) * std::pow(tau, -2.0)));
  set_block(Q, d, d, scale(Qc_inv, 4.0 * std::pow(tau, -1.0)));
  return Q;
}
static Mat calcPhi(int d, double tau) {
  Mat P = Mat::Identity(2 * d);
  for (int i = 0; i < d; i++) P(i, d + i) = tau;
  return P;
}
static Mat calcLambda(const Mat& Qc, double delta_t, double tau) {
  const int d = Qc.r;
  return add(calcPhi(d, tau),
             mul(mul(mul(calcQ(Qc, tau), transpose(calcPhi(d, delta_t - tau))), calcQ_inv(Qc, delta_t)),
                 calcPhi(d, delta_t)),
             -1.0);
}
static Mat calcPsi(const Mat& Qc, double delta_t, double tau) {
  const int d = Qc.r;
  return mul(mul(calcQ(Qc, tau), transpose(calcPhi(d, delta_t - tau))), calcQ_inv(Qc, delta_t));
}
// getQc, gpmp2/gp/GPutils.cpp:16-20: (R^T R)^-1 of the Gaussian model built from the covariance
static Mat getQc(const Mat& Qc_cov) {
  const Mat R = chol_upper(inverse(Qc_cov));
  return inverse(mul(transpose(R), R));
}

// Pose2Vector group helpers (gpmp2/geometry/ProductDynamicLieGroup.h:84-222).  A Pose2Vector is
// stored flat as (x, y, theta, q_1..q_n).
static Vec p2v_between_logmap(int dof, const double* p1, const double* p2, Mat* Hlog_comp1_inv, Mat* Hlog_comp2) {
  // r = Logmap(Compose(Inverse(p1), p2)) with Jacobians Hlogmap*Hcomp1*Hinv and Hlogmap*Hcomp2
  // (GaussianProcessPriorLie.h:71-80).  Inverse: -Ad(p1); Compose(a,b): H1 = Ad(b^-1), H2 = I.
  Pose2 a{p1[0], p1[1], p1[2]}, b{p2[0], p2[1], p2[2]};
  const Pose2 ainv = p2_inverse(a);
  const Pose2 btw = p2_compose(ainv, b);
  Vec r(dof);
  p2_logmap(btw, r.data());
  for (int i = 3; i < dof; i++) r[i] = p2[i] - p1[i];  // vector part: -p1 + p2
  if (Hlog_comp1_inv || Hlog_comp2) {
    const Mat Hlog = p2_logmap_derivative(btw);
    if (Hlog_comp1_inv) {
      Mat Hinv = scale(p2_adjoint(a), -1.0);
      Mat Hcomp1 = p2_adjoint(p2_inverse(b));
      Mat P = mul(mul(Hlog, Hcomp1), Hinv);
      *Hlog_comp1_inv = Mat(dof, dof);
      set_block(*Hlog_comp1_inv, 0, 0, P);
      for (int i = 3; i < dof; i++) (*Hlog_comp1_inv)(i, i) = -1.0;
    }
    if (Hlog_comp2) {
      *Hlog_comp2 = Mat::Identity(dof);
      set_block(*Hlog_comp2, 0, 0, Hlog);
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// GP interpolation
// ------------------------------------------------------------------------------------------------
struct Interp {
  int dof;
  bool lie;
  Mat Lambda, Psi;
  // GaussianProcessInterpolatorLinear ctor, gpmp2/gp/GaussianProcessInterpolatorLinear.h:48-55
  // (same in GaussianProcessInterpolatorLie.h:52-59)
  Interp(const Mat& Qc_cov, double delta_t, double tau, bool lie_) : dof(Qc_cov.r), lie(lie_) {
    const Mat Qc = getQc(Qc_cov);
    Lambda = calcLambda(Qc, delta_t, tau);
    Psi = calcPsi(Qc, delta_t, tau);
  }
  // interpolatePose: Linear.h:62-84 ; Lie.h:64-100.  H[0..3] optional (dof x dof each).
  Vec interpolatePose(const double* p1, const double* v1, const double* p2, const double* v2, Mat* H) const {
    const int d = dof;
    if (!lie) {
      Vec x1(2 * d), x2(2 * d);
      for (int i = 0; i < d; i++) { x1[i] = p1[i]; x1[d + i] = v1[i]; x2[i] = p2[i]; x2[d + i] = v2[i]; }
      if (H) {
        H[0] = block(Lambda, 0, 0, d, d);
        H[1] = block(Lambda, 0, d, d, d);
        H[2] = block(Psi, 0, 0, d, d);
        H[3] = block(Psi, 0, d, d, d);
      }
      const Vec a = mulv(block(Lambda, 0, 0, d, 2 * d), x1), b = mulv(block(Psi, 0, 0, d, 2 * d), x2);
      Vec out(d);
      for (int i = 0; i < d; i++) out[i] = a[i] + b[i];
      return out;
    }
    // Lie version
    Vec r1(2 * d, 0.0), r2(2 * d);
    for (int i = 0; i < d; i++) r1[d + i] = v1[i];
    Mat HlogC1inv, HlogC2;
    const Vec r = p2v_between_logmap(d, p1, p2, H ? &HlogC1inv : nullptr, H ? &HlogC2 : nullptr);
    for (int i = 0; i < d; i++) { r2[i] = r[i]; r2[d + i] = v2[i]; }
    const Vec a = mulv(block(Lambda, 0, 0, d, 2 * d), r1), b = mulv(block(Psi, 0, 0, d, 2 * d), r2);
    Vec xi(d);
    for (int i = 0; i < d; i++) xi[i] = a[i] + b[i];
    // pose = Compose(pose1, Expmap(xi, Hexp), Hcomp21, Hcomp22)
    const Pose2 e = p2_expmap(xi.data());
    const Pose2 a1{p1[0], p1[1], p1[2]};
    const Pose2 pose = p2_compose(a1, e);
    Vec out(d);
    out[0] = pose.x; out[1] = pose.y; out[2] = pose.th;
    for (int i = 3; i < d; i++) out[i] = p1[i] + xi[i];
    if (H) {
      Mat Hexp = Mat::Identity(d);
      set_block(Hexp, 0, 0, p2_expmap_derivative(xi.data()));
      Mat Hcomp21 = Mat::Identity(d);
      set_block(Hcomp21, 0, 0, p2_adjoint(p2_inverse(e)));
      // Hcomp22 = I
      const Mat& Hexpr1 = Hexp;
      const Mat Psi11 = block(Psi, 0, 0, d, d);
      H[0] = add(Hcomp21, mul(mul(Hexpr1, Psi11), HlogC1inv));
      H[1] = mul(Hexpr1, block(Lambda, 0, d, d, d));
      H[2] = mul(mul(Hexpr1, Psi11), HlogC2);
      H[3] = mul(Hexpr1, block(Psi, 0, d, d, d));
    }
    return out;
  }
  // interpolateVelocity: Linear.h:99-123 ; Lie.h:117-148 (no Jacobians needed by the trajectory utilities)
  Vec interpolateVelocity(const double* p1, const double* v1, const double* p2, const double* v2) const {
    const int d = dof;
    Vec r1(2 * d, 0.0), r2(2 * d);
    if (!lie) {
      for (int i = 0; i < d; i++) { r1[i] = p1[i]; r1[d + i] = v1[i]; r2[i] = p2[i]; r2[d + i] = v2[i]; }
    } else {
      for (int i = 0; i < d; i++) r1[d + i] = v1[i];
      const Vec r = p2v_between_logmap(d, p1, p2, nullptr, nullptr);
      for (int i = 0; i < d; i++) { r2[i] = r[i]; r2[d + i] = v2[i]; }
    }
    const Vec a = mulv(block(Lambda, d, 0, d, 2 * d), r1), b = mulv(block(Psi, d, 0, d, 2 * d), r2);
    Vec out(d);
    for (int i = 0; i < d; i++) out[i] = a[i] + b[i];
    return out;
  }
};

// ObstacleSDFFactorGP<ROBOT,GPINTER>::evaluateError, gpmp2/obstacle/ObstacleSDFFactorGP-inl.h:18-75
// (planar twin ObstaclePlanarSDFFactorGP-inl.h:18-78); updatePoseJacobians
// GaussianProcessInterpolatorLinear.h:88-96: H_k = Jerr_conf * Hint_k.
static Vec obstacle_gp_factor(const Robot& rb, const Sdf& f, const Interp& gp, const double* x1, const double* v1,
                              const double* x2, const double* v2, double epsilon, Mat* H /*[4] or NULL*/) {
  Mat Hint[4];
  const Vec conf = gp.interpolatePose(x1, v1, x2, v2, H ? Hint : nullptr);
  Mat Jerr_conf;
  const Vec err = obstacle_factor(rb, f, conf.data(), epsilon, H ? &Jerr_conf : nullptr);
  if (H)
    for (int k = 0; k < 4; k++) H[k] = mul(Jerr_conf, Hint[k]);
  return err;
}

// GaussianProcessPriorLinear::evaluateError, gpmp2/gp/GaussianProcessPriorLinear.h:57-83
// GaussianProcessPriorLie<T>::evaluateError, gpmp2/gp/GaussianProcessPriorLie.h:61-86
static Vec gp_prior_factor(int d, bool lie, double delta_t, const double* p1, const double* v1, const double* p2,
                           const double* v2, Mat* H /*[4] or NULL*/) {
  Vec e(2 * d);
  if (!lie) {
    if (H) {
      H[0] = Mat(2 * d, d); H[1] = Mat(2 * d, d); H[2] = Mat(2 * d, d); H[3] = Mat(2 * d, d);
      for (int i = 0; i < d; i++) {
        H[0](i, i) = 1.0;
        H[1](i, i) = delta_t; H[1](d + i, i) = 1.0;
        H[2](i, i) = -1.0;
        H[3](d + i, i) = -1.0;
      }
    }
    // calcPhi(dof, delta_t) * x1 - x2
    for (int i = 0; i < d; i++) {
      e[i] = (p1[i] + delta_t * v1[i]) - p2[i];
      e[d + i] = v1[i] - v2[i];
    }
    return e;
  }
  Mat HlogC1inv, HlogC2;
  const Vec r = p2v_between_logmap(d, p1, p2, H ? &HlogC1inv : nullptr, H ? &HlogC2 : nullptr);
  if (H) {
    H[0] = Mat(2 * d, d); H[1] = Mat(2 * d, d); H[2] = Mat(2 * d, d); H[3] = Mat(2 * d, d);
    set_block(H[0], 0, 0, HlogC1inv);
    set_block(H[2], 0, 0, HlogC2);
    for (int i = 0; i < d; i++) {
      H[1](i, i) = -delta_t; H[1](d + i, i) = -1.0;
      H[3](d + i, i) = 1.0;
    }
  }
  for (int i = 0; i < d; i++) {
    e[i] = r[i] - v1[i] * delta_t;
    e[d + i] = v2[i] - v1[i];
  }
  return e;
}

// hingeLossJointLimitCost, gpmp2/kinematics/JointLimitCost.h:16-31
static double hinge_limit(double p, double down, double up, double thresh, double* H_p) {
  if (p < down + thresh) {
    if (H_p) *H_p = -1.0;
    return down + thresh - p;
  } else if (p <= up - thresh) {
    if (H_p) *H_p = 0.0;
    return 0.0;
  } else {
    if (H_p) *H_p = 1.0;
    return p - up + thresh;
  }
}

// GoalFactorArm::evaluateError (gpmp2/kinematics/GoalFactorArm.h:52-70) == GaussianPriorWorkspacePosition::evaluateError
// (GaussianPriorWorkspacePosition.h:54-71): e = translation(joint_pos[link]) - goal, H = Hpp * J_jpx_jp[link] with
// Hpp = d translation / d pose = [0 | R]  [GTSAM Pose3::translation(H)].  link < 0 = last joint frame.
static Vec goal_factor(const Robot& rb, const double* conf, int link, const double* goal, Mat* H1) {
  std::vector<M4> px;
  std::vector<Mat> J;
  robot_fk(rb, conf, px, H1 ? &J : nullptr);
  if (link < 0) link = rb.nr_links - 1;
  const M4& T = px[link];
  Vec e(3);
  for (int i = 0; i < 3; i++) e[i] = T[i * 4 + 3] - goal[i];
  if (H1) {
    Mat Hpp(3, 6);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Hpp(i, 3 + j) = T[i * 4 + j];
    *H1 = mul(Hpp, J[link]);
  }
  return e;
}

// Rot3::Logmap of a rotation matrix [GTSAM-recalled 4.0.x, SO3::Logmap]: away from theta = 0 and pi,
// omega = theta / (2 sin theta) * (R32 - R23, R13 - R31, R21 - R12), theta = acos((tr - 1) / 2); near 0 the factor is
// 0.5 - (tr - 3)^2 / 12 [sic]; near pi GTSAM switches to a diagonal-based formula.
static void rot3_logmap(const double R[9], double w[3]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr + 1.0 < 1e-10) {   // theta = pi (+- 2 k pi): pick the axis from the largest diagonal entry
    if (std::fabs(R[8] + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R[8]);
      w[0] = f * R[2]; w[1] = f * R[5]; w[2] = f * (1.0 + R[8]);
    } else if (std::fabs(R[4] + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R[4]);
      w[0] = f * R[1]; w[1] = f * (1.0 + R[4]); w[2] = f * R[7];
    } else {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R[0]);
      w[0] = f * (1.0 + R[0]); w[1] = f * R[3]; w[2] = f * R[6];
    }
    return;
  }
  double magnitude;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-7) {
    const double theta = std::acos((tr - 1.0) / 2.0);
    magnitude = theta / (2.0 * std::sin(theta));
  } else {
    magnitude = 0.5 - tr_3 * tr_3 / 12.0;
  }
  w[0] = magnitude * (R[7] - R[5]); w[1] = magnitude * (R[2] - R[6]); w[2] = magnitude * (R[3] - R[1]);
}
// SO3::LogmapDerivative [GTSAM-recalled]: I + W/2 + (1/theta^2 - (1 + cos theta) / (2 theta sin theta)) W^2, W = skew(omega);
// I + W/2 for theta^2 <= epsilon
static Mat rot3_logmap_derivative(const double w[3]) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  Mat W(3, 3);
  W(0, 1) = -w[2]; W(0, 2) = w[1]; W(1, 0) = w[2]; W(1, 2) = -w[0]; W(2, 0) = -w[1]; W(2, 1) = w[0];
  Mat J = add(Mat::Identity(3), W, 0.5);
  if (t2 <= std::numeric_limits<double>::epsilon()) return J;
  const double t = std::sqrt(t2);
  return add(J, mul(W, W), 1.0 / t2 - (1.0 + std::cos(t)) / (2.0 * t * std::sin(t)));
}
// GaussianPriorWorkspaceOrientation::evaluateError (gpmp2/kinematics/GaussianPriorWorkspaceOrientation.h:53-72):
// e = des.logmap(R) = Logmap(des^T R), H = H_er * H_rp * J_jpx_jp[link] with H_rp = [I 0] (Pose3::rotation) and
// H_er = LogmapDerivative(e) (Rot3::logmap's Jacobian in its argument).  link < 0 = last link frame.
static Vec orientation_factor(const Robot& rb, const double* conf, int link, const double* des /*row-major 3x3*/, Mat* H1) {
  std::vector<M4> px;
  std::vector<Mat> J;
  robot_fk(rb, conf, px, H1 ? &J : nullptr);
  if (link < 0) link = rb.nr_links - 1;
  const M4& T = px[link];
  double E[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double v = 0.0;
      for (int k = 0; k < 3; k++) v += des[k * 3 + i] * T[k * 4 + j];
      E[i * 3 + j] = v;
    }
  Vec e(3);
  rot3_logmap(E, e.data());
  if (H1) *H1 = mul(rot3_logmap_derivative(e.data()), block(J[link], 0, 0, 3, rb.dof));
  return e;
}

// Pose3::Logmap [GTSAM-recalled 4.0.x]: xi = [omega; u], omega = Rot3::Logmap(R); u = t for |omega| < 1e-10, else with
// W = skew(omega / theta): u = t - (theta / 2) W t + (1 - theta / (2 tan(theta / 2))) W W t.
static void pose3_logmap(const double R[9], const double t[3], double xi[6]) {
  rot3_logmap(R, xi);
  const double th = std::sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  if (th < 1e-10) { xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2]; return; }
  const double k[3] = {xi[0] / th, xi[1] / th, xi[2] / th};
  const double Wt[3] = {k[1] * t[2] - k[2] * t[1], k[2] * t[0] - k[0] * t[2], k[0] * t[1] - k[1] * t[0]};
  const double WWt[3] = {k[1] * Wt[2] - k[2] * Wt[1], k[2] * Wt[0] - k[0] * Wt[2], k[0] * Wt[1] - k[1] * Wt[0]};
  const double c = 1.0 - th / (2.0 * std::tan(0.5 * th));
  for (int i = 0; i < 3; i++) xi[3 + i] = t[i] - 0.5 * th * Wt[i] + c * WWt[i];
}
static Mat skew3(const double v[3]) {
  Mat W(3, 3);
  W(0, 1) = -v[2]; W(0, 2) = v[1]; W(1, 0) = v[2]; W(1, 2) = -v[0]; W(2, 0) = -v[1]; W(2, 1) = v[0];
  return W;
}
// Pose3::LogmapDerivative [GTSAM-recalled 4.0.x]: [[Jw, 0], [-Jw Q Jw, Jw]], Jw = Rot3::LogmapDerivative(omega),
// Q = computeQforExpmapDerivative(xi) (Barfoot, State Estimation for Robotics, eq. 7.86b with GTSAM's sign convention).
// Pinned by the analytic-vs-numerical check of testGaussianPriorWorkspacePose.cpp.
static Mat pose3_logmap_derivative(const double xi[6]) {
  const Mat W = skew3(xi), V = skew3(xi + 3);
  const double phi = std::sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  const Mat WV = mul(W, V), VW = mul(V, W), WVW = mul(WV, W), WW = mul(W, W);
  const Mat A1 = add(add(WV, VW), WVW, -1.0);                              // WV + VW - WVW
  const Mat A2 = add(add(mul(WW, V), mul(VW, W)), WVW, -3.0);               // WWV + VWW - 3 WVW
  const Mat A3 = add(mul(WVW, W), mul(W, WVW));                             // WVWW + WWVW
  double c1, c2, c3;
  if (std::fabs(phi) > 1e-5) {
    const double s = std::sin(phi), c = std::cos(phi), p2 = phi * phi, p3 = p2 * phi, p4 = p3 * phi, p5 = p4 * phi;
    c1 = (phi - s) / p3;
    c2 = (1.0 - p2 / 2.0 - c) / p4;
    c3 = -0.5 * ((1.0 - p2 / 2.0 - c) / p4 - 3.0 * (phi - s - p3 / 6.0) / p5);
  } else {
    c1 = 1.0 / 6.0; c2 = 1.0 / 24.0; c3 = -0.5 * (1.0 / 24.0 + 3.0 / 120.0);
  }
  Mat Q = scale(V, -0.5);
  Q = add(Q, A1, c1); Q = add(Q, A2, c2); Q = add(Q, A3, c3);
  const Mat Jw = rot3_logmap_derivative(xi);
  const Mat Q2 = scale(mul(mul(Jw, Q), Jw), -1.0);
  Mat J(6, 6);
  set_block(J, 0, 0, Jw); set_block(J, 3, 0, Q2); set_block(J, 3, 3, Jw);
  return J;
}
// GaussianPriorWorkspacePose::evaluateError (gpmp2/kinematics/GaussianPriorWorkspacePose.h:53-70):
// e = des.logmap(T_link) = Logmap(des^-1 T_link), H = LogmapDerivative(des^-1 T_link) * J_jpx_jp[link].
// des: row-major 3x3 rotation + translation.  link < 0 = last link frame.
static Vec pose_factor(const Robot& rb, const double* conf, int link, const double* desR, const double* dest, Mat* H1) {
  std::vector<M4> px;
  std::vector<Mat> J;
  robot_fk(rb, conf, px, H1 ? &J : nullptr);
  if (link < 0) link = rb.nr_links - 1;
  const M4& T = px[link];
  double E[9], te[3];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double v = 0.0;
      for (int k = 0; k < 3; k++) v += desR[k * 3 + i] * T[k * 4 + j];
      E[i * 3 + j] = v;
    }
    double v = 0.0;
    for (int k = 0; k < 3; k++) v += desR[k * 3 + i] * (T[k * 4 + 3] - dest[k]);
    te[i] = v;
  }
  Vec e(6);
  pose3_logmap(E, te, e.data());
  if (H1) *H1 = mul(pose3_logmap_derivative(e.data()), J[link]);
  return e;
}

// SelfCollision<ROBOT>::evaluateError + hingeLossSelfCollisionCost (gpmp2/obstacle/SelfCollision.h:66-128):
// data rows (sphere A id, sphere B id, epsilon, sigma); e_p = hinge(r_A + r_B + epsilon - distance3(c_A, c_B)),
// H row = [-H_A, -H_B] * [J_A; J_B] with distance3 Jacobians H_A = (c_A - c_B)^T / dist = -H_B  [GTSAM distance3].
static Vec self_collision_factor(const Robot& rb, const double* conf, int n, const double* data, Mat* H1) {
  Vec centers;
  std::vector<Mat> J;
  sphere_centers(rb, conf, centers, H1 ? &J : nullptr);
  Vec e(n, 0.0);
  if (H1) *H1 = Mat(n, rb.dof);
  for (int p = 0; p < n; p++) {
    const int A = (int)data[4 * p + 0], B = (int)data[4 * p + 1];
    const double total_eps = rb.sph_radius[A] + rb.sph_radius[B] + data[4 * p + 2];
    double d[3], dist = 0.0;
    for (int k = 0; k < 3; k++) { d[k] = centers[3 * A + k] - centers[3 * B + k]; dist += d[k] * d[k]; }
    dist = std::sqrt(dist);
    if (dist > total_eps) continue;   // SelfCollision.h:111-117: zero error, zero Jacobian
    e[p] = total_eps - dist;
    if (H1)
      for (int c = 0; c < rb.dof; c++) {
        double v = 0.0;
        for (int k = 0; k < 3; k++) v += -(d[k] / dist) * J[A](k, c) + (d[k] / dist) * J[B](k, c);
        (*H1)(p, c) = v;
      }
  }
  return e;
}

// ------------------------------------------------------------------------------------------------
// The planning problem = the factor graph of internal::BatchTrajOptimize,
// gpmp2/planner/BatchTrajOptimizer-inl.h:19-84.
// Variables: var 2i = x_i, var 2i+1 = v_i, each of tangent dim D; flat state as in gpmp2b.h.
// ------------------------------------------------------------------------------------------------
struct LinFactor {          // a whitened Jacobian factor  || sum_k A_k d_k - b ||^2
  int nvars;
  int vars[4];
  Mat A[4];
  Vec b;
};

struct Problem {
  const Robot& rb;
  const Sdf& sdf;
  const gpmp2b_setting& st;
  int D, N, K;
  bool lie;
  double delta_t, inter_dt;
  Mat Qc_cov, R_gp;              // R_gp^T R_gp = Q(delta_t)^-1
  std::vector<Interp> interps;   // one per j=1..K (the reference builds one per factor object)
  const double *start_conf, *start_vel, *end_conf, *end_vel;
  const double *fix_conf = nullptr, *fix_vel = nullptr;   // this problem's rows of gpmp2b_setting.fix_conf / fix_vel

  Problem(const Robot& r, const Sdf& s, const gpmp2b_setting& set, const double* sc, const double* sv,
          const double* ec, const double* ev)
      : rb(r), sdf(s), st(set), start_conf(sc), start_vel(sv), end_conf(ec), end_vel(ev) {
    D = st.dof;
    N = st.total_step + 1;
    K = st.obs_check_inter;
    lie = rb.kind != GPMP2B_ROBOT_ARM;
    delta_t = st.total_time / static_cast<double>(st.total_step);          // -inl.h:30
    inter_dt = delta_t / static_cast<double>(st.obs_check_inter + 1);      // -inl.h:31
    Qc_cov = Mat::Identity(D);
    if (st.Qc)
      for (int i = 0; i < D * D; i++) Qc_cov.a[i] = st.Qc[i];
    // GaussianProcessPriorLinear noise: Gaussian::Covariance(calcQ(getQc(Qc_model), delta_t))
    // (GaussianProcessPriorLinear.h:43) -> R = chol_upper(Q^-1)  [GTSAM Gaussian::Information]
    R_gp = chol_upper(inverse(calcQ(getQc(Qc_cov), delta_t)));
    for (int j = 1; j <= K; j++) interps.emplace_back(Qc_cov, delta_t, inter_dt * static_cast<double>(j), lie);
  }
  const double* X(const Vec& t, int i) const { return &t[(size_t)i * D]; }
  const double* V(const Vec& t, int i) const { return &t[(size_t)(N + i) * D]; }

  // PriorFactor<T>::evaluateError = -Local(x, prior), H = I  [GTSAM-recalled, SURVEY App. B.3].
  // Pose2Vector Local: Pose2 chart (between.x, between.y, between.theta) + vector difference
  // (ProductDynamicLieGroup.h:92-100).
  Vec prior_pose_err(const double* x, const double* prior) const {
    Vec e(D);
    if (!lie) {
      for (int i = 0; i < D; i++) e[i] = x[i] - prior[i];
      return e;
    }
    const Pose2 a{x[0], x[1], x[2]}, b{prior[0], prior[1], prior[2]};
    const Pose2 btw = p2_compose(p2_inverse(a), b);
    e[0] = -btw.x; e[1] = -btw.y; e[2] = -btw.th;
    for (int i = 3; i < D; i++) e[i] = -(prior[i] - x[i]);
    return e;
  }

  // Enumerate every factor; `fn(nvars, vars, err(unwhitened), H blocks or null, whiten)` does the
  // work.  Order follows BatchTrajOptimizer-inl.h:36-81.
  template <class F>
  void for_each_factor(const Vec& t, bool want_H, F&& fn) const {
    const int S = rb.nr_spheres();
    for (int i = 0; i < N; i++) {
      const int xk = 2 * i, vk = 2 * i + 1;
      if (i == 0 || i == N - 1) {  // -inl.h:41-48
        const double* pc = i == 0 ? start_conf : end_conf;
        const double* pv = i == 0 ? start_vel : end_vel;
        // optional workspace goal (gpmp2b_setting.goal_*): the hand-built graph of matlab/Arm3GoalReachExample.m:104-108
        // puts GoalFactorArm on x_T instead of the end-configuration prior
        if (i == N - 1 && st.goal_enabled) {
          Mat H[1];
          Vec e = st.goal_enabled == 2 ? pose_factor(rb, X(t, i), st.goal_link, st.goal_R, st.goal_pos, want_H ? &H[0] : nullptr)
                                       : goal_factor(rb, X(t, i), st.goal_link, st.goal_pos, want_H ? &H[0] : nullptr);
          int vars[1] = {xk};
          fn(1, vars, e, want_H ? H : nullptr, 0 /*iso*/, st.goal_sigma, (const double*)nullptr);
        }
        if (!(i == N - 1 && st.goal_enabled && !st.goal_keep_end_prior)) {
          Vec e = prior_pose_err(X(t, i), pc);
          Mat H[1] = {Mat::Identity(D)};
          int vars[1] = {xk};
          fn(1, vars, e, want_H ? H : nullptr, 0 /*iso*/, st.conf_prior_sigma, (const double*)nullptr);
        }
        {
          Vec e(D);
          for (int k = 0; k < D; k++) e[k] = V(t, i)[k] - pv[k];
          Mat H[1] = {Mat::Identity(D)};
          int vars[1] = {vk};
          fn(1, vars, e, want_H ? H : nullptr, 0, st.vel_prior_sigma, (const double*)nullptr);
        }
      }
      // optional pinned state of a replanning re-solve (gpmp2b_setting.fix_*): ISAM2TrajOptimizer::fixConfigAndVel,
      // gpmp2/planner/ISAM2TrajOptimizer-inl.h:160-168 -- PriorFactor<Pose>(x_k, conf_fix, conf_prior_model) and
      // PriorFactor<Velocity>(v_k, vel_fix, vel_prior_model)
      if (st.fix_enabled && i == st.fix_state_index) {
        {
          Vec e = prior_pose_err(X(t, i), fix_conf);
          Mat H[1] = {Mat::Identity(D)};
          int vars[1] = {xk};
          fn(1, vars, e, want_H ? H : nullptr, 0 /*iso*/, st.conf_prior_sigma, (const double*)nullptr);
        }
        {
          Vec e(D);
          for (int k = 0; k < D; k++) e[k] = V(t, i)[k] - fix_vel[k];
          Mat H[1] = {Mat::Identity(D)};
          int vars[1] = {vk};
          fn(1, vars, e, want_H ? H : nullptr, 0, st.vel_prior_sigma, (const double*)nullptr);
        }
      }
      if (st.flag_pos_limit) {  // -inl.h:50-54, JointLimitFactorVector.h:62-79 / ...Pose2Vector.h:66-91
        Vec e(D, 0.0);
        Mat H[1] = {Mat(D, D)};
        const int off = lie ? 3 : 0;
        for (int k = off; k < D; k++) {
          double Hp;
          e[k] = hinge_limit(X(t, i)[k], st.joint_pos_limits_down[k], st.joint_pos_limits_up[k],
                             st.pos_limit_thresh[k], &Hp);
          H[0](k, k) = Hp;
        }
        int vars[1] = {xk};
        fn(1, vars, e, want_H ? H : nullptr, 1 /*diag*/, 0.0, st.pos_limit_sigma);
      }
      if (st.flag_vel_limit) {  // -inl.h:55-59, VelocityLimitFactorVector.h:62-79
        Vec e(D, 0.0);
        Mat H[1] = {Mat(D, D)};
        for (int k = 0; k < D; k++) {
          double Hp;
          e[k] = hinge_limit(V(t, i)[k], -st.vel_limits[k], st.vel_limits[k], st.vel_limit_thresh[k], &Hp);
          H[0](k, k) = Hp;
        }
        int vars[1] = {vk};
        fn(1, vars, e, want_H ? H : nullptr, 1, 0.0, st.vel_limit_sigma);
      }
      if (st.n_self_collision > 0) {  // optional SelfCollisionArm on every support state (gpmp2b_setting), Diagonal::Sigmas(data.col(3))
        const int n = st.n_self_collision;
        Mat H[1];
        Vec e = self_collision_factor(rb, X(t, i), n, st.self_collision_data, want_H ? &H[0] : nullptr);
        Vec sig(n);
        for (int p = 0; p < n; p++) sig[p] = st.self_collision_data[4 * p + 3];
        int vars[1] = {xk};
        fn(1, vars, e, want_H ? H : nullptr, 1 /*diag*/, 0.0, sig.data());
      }
      if (st.orient_enabled && i >= st.orient_state_first && i <= st.orient_state_last) {  // optional GaussianPriorWorkspaceOrientation (gpmp2b_setting)
        Mat H[1];
        Vec e = orientation_factor(rb, X(t, i), st.orient_link, st.orient_R, want_H ? &H[0] : nullptr);
        int vars[1] = {xk};
        fn(1, vars, e, want_H ? H : nullptr, 0 /*iso*/, st.orient_sigma, (const double*)nullptr);
      }
      if (st.vehicle_dynamics_sigma > 0.0) {  // optional VehicleDynamicsFactorPose2Vector on every support state (gpmp2b_setting):
        // simple2DVehicleDynamicsPose2 (dynamics/VehicleDynamics.h:19-28): e = v(1), Hp = 0, Hv = (0, 1, 0);
        // VehicleDynamicsFactorPose2Vector::evaluateError (VehicleDynamicsFactorPose2Vector.h:56-79) pads both to 1 x D
        Vec e(1, V(t, i)[1]);
        Mat H[2] = {Mat(1, D), Mat(1, D)};
        H[1](0, 1) = 1.0;
        int vars[2] = {xk, vk};
        fn(2, vars, e, want_H ? H : nullptr, 0 /*iso*/, st.vehicle_dynamics_sigma, (const double*)nullptr);
      }
      {  // -inl.h:62 unary obstacle factor, Isotropic(S, cost_sigma) (ObstacleSDFFactor.h:67)
        Mat H[1];
        Vec e = obstacle_factor(rb, sdf, X(t, i), st.epsilon, want_H ? &H[0] : nullptr);
        int vars[1] = {xk};
        fn(1, vars, e, want_H ? H : nullptr, 0, st.cost_sigma, (const double*)nullptr);
      }
      if (i > 0) {
        int vars[4] = {2 * (i - 1), 2 * (i - 1) + 1, xk, vk};
        for (int j = 1; j <= K; j++) {  // -inl.h:69-75
          Mat H[4];
          Vec e = obstacle_gp_factor(rb, sdf, interps[j - 1], X(t, i - 1), V(t, i - 1), X(t, i), V(t, i),
                                     st.epsilon, want_H ? H : nullptr);
          fn(4, vars, e, want_H ? H : nullptr, 0, st.cost_sigma, (const double*)nullptr);
        }
        {  // -inl.h:78-79 GP prior
          Mat H[4];
          Vec e = gp_prior_factor(D, lie, delta_t, X(t, i - 1), V(t, i - 1), X(t, i), V(t, i), want_H ? H : nullptr);
          fn(4, vars, e, want_H ? H : nullptr, 2 /*full gaussian R_gp*/, 0.0, (const double*)nullptr);
        }
      }
    }
    (void)S;
  }

  // whiten in place: kind 0 Isotropic(sigma), 1 Diagonal::Sigmas(sig), 2 Gaussian R_gp
  void whiten(int kind, double sigma, const double* sig, Vec& e, Mat* H, int nH) const {
    if (kind == 0) {
      for (auto& x : e) x /= sigma;
      if (H) for (int k = 0; k < nH; k++) for (auto& x : H[k].a) x /= sigma;
    } else if (kind == 1) {
      for (size_t r = 0; r < e.size(); r++) e[r] /= sig[r];
      if (H) for (int k = 0; k < nH; k++) for (int r = 0; r < H[k].r; r++) for (int c = 0; c < H[k].c; c++) H[k](r, c) /= sig[r];
    } else {
      e = mulv(R_gp, e);
      if (H) for (int k = 0; k < nH; k++) H[k] = mul(R_gp, H[k]);
    }
  }

  // NonlinearFactorGraph::error = 0.5 * sum ||whitened||^2
  double error(const Vec& t) const {
    double total = 0.0;
    for_each_factor(t, false, [&](int nv, const int*, Vec& e, Mat*, int kind, double sigma, const double* sig) {
      (void)nv;
      whiten(kind, sigma, sig, e, nullptr, 0);
      double s = 0;
      for (double x : e) s += x * x;
      total += 0.5 * s;
    });
    return total;
  }

  // NonlinearFactorGraph::linearize: whitened (A, b = -e) per factor
  void linearize(const Vec& t, std::vector<LinFactor>& out) const {
    out.clear();
    for_each_factor(t, true, [&](int nv, const int* vars, Vec& e, Mat* H, int kind, double sigma, const double* sig) {
      whiten(kind, sigma, sig, e, H, nv);
      LinFactor lf;
      lf.nvars = nv;
      for (int k = 0; k < nv; k++) { lf.vars[k] = vars[k]; lf.A[k] = H[k]; }
      lf.b.resize(e.size());
      for (size_t r = 0; r < e.size(); r++) lf.b[r] = -e[r];
      out.push_back(std::move(lf));
    });
  }

  // Values::retract: vectors x+d; Pose2Vector: Pose2 chart retract p * Pose2(dx,dy,dth) + vector add
  // (ProductDynamicLieGroup.h:84-90; GTSAM Pose2 default chart)
  Vec retract(const Vec& t, const Vec& delta) const {
    Vec o(t.size());
    for (int i = 0; i < N; i++) {
      const double* dx = &delta[(size_t)(2 * i) * D];
      const double* dv = &delta[(size_t)(2 * i + 1) * D];
      const double* x = X(t, i);
      double* ox = &o[(size_t)i * D];
      double* ov = &o[(size_t)(N + i) * D];
      if (!lie) {
        for (int k = 0; k < D; k++) ox[k] = x[k] + dx[k];
      } else {
        const Pose2 p = p2_compose(Pose2{x[0], x[1], x[2]}, Pose2{dx[0], dx[1], dx[2]});
        ox[0] = p.x; ox[1] = p.y; ox[2] = p.th;
        for (int k = 3; k < D; k++) ox[k] = x[k] + dx[k];
      }
      for (int k = 0; k < D; k++) ov[k] = V(t, i)[k] + dv[k];
    }
    return o;
  }
};

// dense normal equations in variable order [x_0, v_0, x_1, v_1, ...]: H = sum A^T A, g = -sum A^T b
static void normal_equations(const std::vector<LinFactor>& lin, int nvars, int D, Mat& H, Vec& g) {
  const int n = nvars * D;
  H = Mat(n, n);
  g.assign(n, 0.0);
  for (const LinFactor& f : lin) {
    const int m = (int)f.b.size();
    for (int ka = 0; ka < f.nvars; ka++) {
      const int ra = f.vars[ka] * D;
      for (int c = 0; c < D; c++) {
        double s = 0;
        for (int r = 0; r < m; r++) s += f.A[ka](r, c) * f.b[r];
        g[ra + c] -= s;
      }
      for (int kb = 0; kb < f.nvars; kb++) {
        const int rb_ = f.vars[kb] * D;
        for (int r = 0; r < m; r++)
          for (int c1 = 0; c1 < D; c1++) {
            const double a1 = f.A[ka](r, c1);
            if (a1 == 0.0) continue;
            for (int c2 = 0; c2 < D; c2++) H(ra + c1, rb_ + c2) += a1 * f.A[kb](r, c2);
          }
      }
    }
  }
}

// GaussianFactorGraph::error(delta) = 0.5 sum ||A d - b||^2
static double linear_error(const std::vector<LinFactor>& lin, int D, const Vec& delta) {
  double total = 0;
  for (const LinFactor& f : lin) {
    const int m = (int)f.b.size();
    double s = 0;
    for (int r = 0; r < m; r++) {
      double v = -f.b[r];
      for (int k = 0; k < f.nvars; k++) {
        const double* d = &delta[(size_t)f.vars[k] * D];
        for (int c = 0; c < D; c++) v += f.A[k](r, c) * d[c];
      }
      s += v * v;
    }
    total += 0.5 * s;
  }
  return total;
}

// Solve (H + lambda I) x = -g by Cholesky restricted to half-bandwidth w (w >= n-1: dense).
// Stands in for GTSAM's multifrontal Cholesky elimination; returns false on a non-positive pivot
// (IndeterminantLinearSystemException).
static bool solve_spd(const Mat& H, const Vec& g, double lambda, int w, Vec& x) {
  const int n = H.r;
  Mat L(n, n);
  for (int j = 0; j < n; j++) {
    const int k0 = std::max(0, j - w);
    double s = H(j, j) + lambda;
    for (int k = k0; k < j; k++) s -= L(j, k) * L(j, k);
    if (!(s > 0.0)) return false;
    const double ljj = std::sqrt(s);
    L(j, j) = ljj;
    const int i1 = std::min(n - 1, j + w);
    for (int i = j + 1; i <= i1; i++) {
      double tsum = H(i, j);
      for (int k = std::max(k0, i - w); k < j; k++) tsum -= L(i, k) * L(j, k);
      L(i, j) = tsum / ljj;
    }
  }
  Vec y(n);
  for (int i = 0; i < n; i++) {
    double s = -g[i];
    for (int k = std::max(0, i - w); k < i; k++) s -= L(i, k) * y[k];
    y[i] = s / L(i, i);
  }
  x.assign(n, 0.0);
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k <= std::min(n - 1, i + w); k++) s -= L(k, i) * x[k];
    x[i] = s / L(i, i);
  }
  return true;
}

// checkConvergence [GTSAM NonlinearOptimizer.cpp, recalled -- SURVEY App. B.2]
static bool check_convergence(double relTol, double absTol, double errTol, double cur, double nw, int* why) {
  if (nw <= errTol) { if (why) *why = GPMP2B_ST_ERROR_TOL; return true; }
  const double absoluteDecrease = cur - nw;
  const double relativeDecrease = absoluteDecrease / cur;
  const bool rel = relTol && (relativeDecrease <= relTol);
  const bool ab = absoluteDecrease <= absTol;
  if (why) *why = (rel ? GPMP2B_ST_CONVERGED_REL : 0) | (ab ? GPMP2B_ST_CONVERGED_ABS : 0);
  return rel || ab;
}

struct OptResult {
  Vec traj;
  double error = 0;
  int iters = 0, status = 0;
  long n_lin = 0, n_solve = 0, n_err = 0;
};

// gpmp2::optimize, gpmp2/planner/BatchTrajOptimizer.cpp:212-308, over the GTSAM optimizers
// [LevenbergMarquardtOptimizer::iterate/tryLambda and GaussNewtonOptimizer::iterate, recalled --
// SURVEY App. B.1/B.3].  `dense` selects the dense Cholesky (tests) or the banded one (timing).
static OptResult optimize(const Problem& P, const Vec& init, bool dense) {
  const gpmp2b_setting& st = P.st;
  const int D = P.D, nvars = 2 * P.N, n = nvars * D;
  const int w = dense ? n : (4 * D - 1);
  // LevenbergMarquardtParams defaults, lambdaInitial overridden at BatchTrajOptimizer.cpp:226
  double lambda = 100.0;
  const double lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3;
  const double absoluteErrorTol = 1e-5, errorTol = 0.0, relativeErrorTol = st.rel_thresh;
  const int maxIterations = st.max_iter;
  // DoglegParams: setDeltaInitial(0.2) (BatchTrajOptimizer.cpp:219-222); DoglegOptimizer uses
  // DoglegOptimizerImpl::Iterate in ONE_STEP_PER_ITERATION mode [GTSAM 4.0.x, recalled]
  double Delta = 0.2;

  OptResult R;
  Vec values = init;
  double error = P.error(values);  // optimizer constructor computes graph.error(initial)
  R.n_err++;
  int iterations = 0;
  double currentError = error;
  auto finish = [&](const Vec& v, double e) { R.traj = v; R.error = e; R.iters = iterations; return R; };
  if (currentError <= errorTol) { R.status |= GPMP2B_ST_ERROR_TOL; return finish(values, error); }
  if (iterations >= maxIterations) { R.status |= GPMP2B_ST_MAX_ITER; return finish(values, error); }

  Vec last_values;
  double last_error = error;
  std::vector<LinFactor> lin;
  Mat H;
  Vec g, delta;
  int why = 0;
  do {
    currentError = error;
    last_values = values;  // iter_no_increase is always true (BatchTrajOptimizer-inl.h:83)
    last_error = error;
    // ---- opt->iterate() ----
    P.linearize(values, lin);
    R.n_lin++;
    normal_equations(lin, nvars, D, H, g);
    if (st.opt_type == GPMP2B_OPT_DOGLEG) {
      // dx_u = optimizeGradientSearch(): -(g.g / g.H.g) g ;  dx_n = optimize(): -H^-1 g
      R.n_solve++;
      Vec dx_n;
      if (!solve_spd(H, g, 0.0, w, dx_n)) { R.status |= GPMP2B_ST_SOLVE_FAILED; break; }
      double gg = 0, gHg = 0;
      for (int i = 0; i < n; i++) {
        gg += g[i] * g[i];
        double hgi = 0;
        for (int j = std::max(0, i - w); j <= std::min(n - 1, i + w); j++) hgi += H(i, j) * g[j];
        gHg += g[i] * hgi;
      }
      const double step = -gg / gHg;
      Vec dx_u(n);
      for (int i = 0; i < n; i++) dx_u[i] = step * g[i];
      const double f_error = error;
      bool stay = true;
      Vec dx_d(n);
      double new_f = f_error;
      while (stay) {
        // ComputeDoglegPoint
        double uu = 0, nn = 0, un = 0;
        for (int i = 0; i < n; i++) { uu += dx_u[i] * dx_u[i]; nn += dx_n[i] * dx_n[i]; un += dx_u[i] * dx_n[i]; }
        const double DeltaSq = Delta * Delta;
        if (DeltaSq < uu) {
          const double sc = std::sqrt(DeltaSq / uu);
          for (int i = 0; i < n; i++) dx_d[i] = sc * dx_u[i];
        } else if (DeltaSq < nn) {   // ComputeBlend
          const double a = uu - 2. * un + nn, b = 2. * (un - uu), c = uu - DeltaSq;
          const double sq = std::sqrt(b * b - 4 * a * c);
          const double tau1 = (-b + sq) / (2. * a), tau2 = (-b - sq) / (2. * a);
          const double tau = (0.0 <= tau1 && tau1 <= 1.0) ? tau1 : tau2;
          for (int i = 0; i < n; i++) dx_d[i] = (1. - tau) * dx_u[i] + tau * dx_n[i];
        } else {
          dx_d = dx_n;
        }
        const Vec x0_plus_dx = P.retract(values, dx_d);
        new_f = P.error(x0_plus_dx);
        R.n_err++;
        const double new_M_error = linear_error(lin, D, dx_d);
        const double M_error = f_error;   // the linear model's error at zero equals the nonlinear error
        const double rho = (std::fabs(f_error - new_f) < 1e-15 || std::fabs(M_error - new_M_error) < 1e-15)
                               ? 0.5 : (f_error - new_f) / (M_error - new_M_error);
        if (rho >= 0.75) {
          double nd = 0;
          for (int i = 0; i < n; i++) nd += dx_d[i] * dx_d[i];
          Delta = std::max(Delta, 3.0 * std::sqrt(nd));
          stay = false;
        } else if (rho >= 0.25) {
          stay = false;
        } else if (rho >= 0.0) {
          if (Delta > 1e-5) Delta = 0.5 * Delta;
          stay = false;   // ONE_STEP_PER_ITERATION
        } else {
          if (Delta > 1e-5) { Delta *= 0.5; stay = true; }
          else { std::fill(dx_d.begin(), dx_d.end(), 0.0); new_f = f_error; stay = false; }
        }
      }
      values = P.retract(values, dx_d);
      error = new_f;
      iterations++;
    } else if (st.opt_type == GPMP2B_OPT_GAUSS_NEWTON) {
      R.n_solve++;
      if (!solve_spd(H, g, 0.0, w, delta)) { R.status |= GPMP2B_ST_SOLVE_FAILED; break; }
      values = P.retract(values, delta);
      error = P.error(values);
      R.n_err++;
      iterations++;
    } else {
      for (;;) {  // while (!tryLambda(...))
        bool step_is_successful = false, stopSearchingLambda = false;
        double newError = std::numeric_limits<double>::infinity();
        Vec newValues;
        R.n_solve++;
        const bool solved = solve_spd(H, g, lambda, w, delta);
        if (solved) {
          const double newlinearizedError = linear_error(lin, D, delta);
          const double linearizedCostChange = error - newlinearizedError;
          if (linearizedCostChange >= 0) {
            newValues = P.retract(values, delta);
            newError = P.error(newValues);
            R.n_err++;
            const double costChange = error - newError;
            if (linearizedCostChange > std::numeric_limits<double>::epsilon() * std::fabs(error)) {
              const double modelFidelity = costChange / linearizedCostChange;
              step_is_successful = modelFidelity > minModelFidelity;
            }
            const double minAbsoluteTolerance = relativeErrorTol * error;
            if (std::fabs(costChange) < minAbsoluteTolerance) stopSearchingLambda = true;
          }
        } else {
          R.status |= GPMP2B_ST_SOLVE_FAILED;
        }
        if (step_is_successful) {
          values = newValues;
          error = newError;
          lambda = std::max(lambdaLowerBound, lambda / lambdaFactor);
          iterations++;
          break;
        } else if (!stopSearchingLambda) {
          lambda *= lambdaFactor;
          if (lambda >= lambdaUpperBound) { R.status |= GPMP2B_ST_LAMBDA_MAXED; break; }
        } else {
          break;
        }
      }
    }
  } while (iterations < maxIterations &&
           !check_convergence(relativeErrorTol, absoluteErrorTol, errorTol, currentError, error, &why));
  if (iterations >= maxIterations) R.status |= GPMP2B_ST_MAX_ITER;
  else R.status |= why;
  // BatchTrajOptimizer.cpp:297-307
  if (error > currentError) {
    R.status |= GPMP2B_ST_ERR_INCREASED;
    return finish(last_values, last_error);
  }
  return finish(values, error);
}

// internal::CollisionCost, gpmp2/planner/BatchTrajOptimizer-inl.h:87-100
static double collision_cost(const Robot& rb, const Sdf& sdf, int D, int N, const double* traj) {
  double c = 0;
  for (int i = 0; i < N; i++) {
    const Vec e = obstacle_factor(rb, sdf, traj + (size_t)i * D, 0.0, nullptr);
    for (double x : e) c += x;
  }
  return c;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C entry points (ctypes / bench.py).  Same structs as the product header.
// ------------------------------------------------------------------------------------------------
using namespace orc;

#define ORC_TRY try {
#define ORC_CATCH                                             \
  }                                                           \
  catch (const std::exception& e) {                           \
    std::fprintf(stderr, "[oracle] %s\n", e.what());          \
    return -1;                                                \
  }                                                           \
  return 0;

extern "C" {

// link poses [nr_links][16] row-major 4x4; J [nr_links][6][dof] (may be NULL)
int orc_forward_kinematics(const gpmp2b_robot_desc* rd, const double* conf, double* out_poses, double* out_J) {
  ORC_TRY
  Robot rb(*rd);
  std::vector<M4> px;
  std::vector<Mat> J;
  robot_fk(rb, conf, px, out_J ? &J : nullptr);
  for (int i = 0; i < rb.nr_links; i++) {
    std::memcpy(out_poses + 16 * i, px[i].data(), 16 * sizeof(double));
    if (out_J) std::memcpy(out_J + (size_t)i * 6 * rb.dof, J[i].a.data(), sizeof(double) * 6 * rb.dof);
  }
  ORC_CATCH
}

// centers [S][3]; J [S][3][dof] (may be NULL)
int orc_sphere_centers(const gpmp2b_robot_desc* rd, const double* conf, double* out_centers, double* out_J) {
  ORC_TRY
  Robot rb(*rd);
  Vec c;
  std::vector<Mat> J;
  sphere_centers(rb, conf, c, out_J ? &J : nullptr);
  std::memcpy(out_centers, c.data(), sizeof(double) * c.size());
  if (out_J)
    for (int s = 0; s < rb.nr_spheres(); s++)
      std::memcpy(out_J + (size_t)s * 3 * rb.dof, J[s].a.data(), sizeof(double) * 3 * rb.dof);
  ORC_CATCH
}

// returns 1 in range, 0 out of range (SDFQueryOutOfRange), -1 error
int orc_sdf_query(const gpmp2b_sdf_desc* sd, const double* p, double* out_dist, double* out_grad) {
  Sdf f(*sd);
  double g[3] = {0, 0, 0}, d = 0;
  const bool ok = f.ndim == 3 ? sdf3_query(f, p, d, g) : sdf2_query(f, p, d, g);
  if (!ok) return 0;
  *out_dist = d;
  for (int i = 0; i < f.ndim; i++) out_grad[i] = g[i];
  return 1;
}

// unary obstacle factor: err [S], H [S][dof] (may be NULL)
int orc_goal_factor(const gpmp2b_robot_desc* rd, const double* conf, int link, const double* goal, double* out_err,
                    double* out_J) {
  Robot rb(*rd);
  Mat H;
  Vec e = goal_factor(rb, conf, link, goal, out_J ? &H : nullptr);
  for (int i = 0; i < 3; i++) out_err[i] = e[i];
  if (out_J) std::memcpy(out_J, H.a.data(), sizeof(double) * 3 * rb.dof);
  return 0;
}

int orc_orientation_factor(const gpmp2b_robot_desc* rd, const double* conf, int link, const double* des, double* out_err,
                           double* out_J) {
  Robot rb(*rd);
  Mat H;
  Vec e = orientation_factor(rb, conf, link, des, out_J ? &H : nullptr);
  for (int i = 0; i < 3; i++) out_err[i] = e[i];
  if (out_J) std::memcpy(out_J, H.a.data(), sizeof(double) * 3 * rb.dof);
  return 0;
}

int orc_pose_factor(const gpmp2b_robot_desc* rd, const double* conf, int link, const double* desR, const double* dest,
                    double* out_err, double* out_J) {
  Robot rb(*rd);
  Mat H;
  Vec e = pose_factor(rb, conf, link, desR, dest, out_J ? &H : nullptr);
  for (int i = 0; i < 6; i++) out_err[i] = e[i];
  if (out_J) std::memcpy(out_J, H.a.data(), sizeof(double) * 6 * rb.dof);
  return 0;
}

int orc_self_collision_factor(const gpmp2b_robot_desc* rd, const double* conf, int n, const double* data, double* out_err,
                              double* out_J) {
  Robot rb(*rd);
  Mat H;
  Vec e = self_collision_factor(rb, conf, n, data, out_J ? &H : nullptr);
  for (int i = 0; i < n; i++) out_err[i] = e[i];
  if (out_J) std::memcpy(out_J, H.a.data(), sizeof(double) * n * rb.dof);
  return 0;
}

int orc_obstacle_factor(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const double* conf, double epsilon,
                        double* out_err, double* out_H) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  Mat H;
  const Vec e = obstacle_factor(rb, f, conf, epsilon, out_H ? &H : nullptr);
  std::memcpy(out_err, e.data(), sizeof(double) * e.size());
  if (out_H) std::memcpy(out_H, H.a.data(), sizeof(double) * H.a.size());
  ORC_CATCH
}

// GP obstacle factor: err [S], H1..H4 [S][dof] each (out_H [4][S][dof], may be NULL); Qc [dof*dof] or NULL
int orc_obstacle_gp_factor(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const double* Qc, double delta_t,
                           double tau, const double* x1, const double* v1, const double* x2, const double* v2,
                           double epsilon, double* out_err, double* out_H) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  Mat Qm = Mat::Identity(rb.dof);
  if (Qc) for (int i = 0; i < rb.dof * rb.dof; i++) Qm.a[i] = Qc[i];
  Interp gp(Qm, delta_t, tau, rb.kind != GPMP2B_ROBOT_ARM);
  Mat H[4];
  const Vec e = obstacle_gp_factor(rb, f, gp, x1, v1, x2, v2, epsilon, out_H ? H : nullptr);
  std::memcpy(out_err, e.data(), sizeof(double) * e.size());
  if (out_H)
    for (int k = 0; k < 4; k++) std::memcpy(out_H + (size_t)k * H[k].a.size(), H[k].a.data(), sizeof(double) * H[k].a.size());
  ORC_CATCH
}

// GP interpolation of the pose: out_pose [dof], out_H [4][dof][dof] (may be NULL). lie: 0 vector, 1 Pose2Vector
int orc_gp_interpolate(int dof, int lie, const double* Qc, double delta_t, double tau, const double* x1,
                       const double* v1, const double* x2, const double* v2, double* out_pose, double* out_H) {
  ORC_TRY
  Mat Qm = Mat::Identity(dof);
  if (Qc) for (int i = 0; i < dof * dof; i++) Qm.a[i] = Qc[i];
  Interp gp(Qm, delta_t, tau, lie != 0);
  Mat H[4];
  const Vec p = gp.interpolatePose(x1, v1, x2, v2, out_H ? H : nullptr);
  std::memcpy(out_pose, p.data(), sizeof(double) * dof);
  if (out_H)
    for (int k = 0; k < 4; k++) std::memcpy(out_H + (size_t)k * dof * dof, H[k].a.data(), sizeof(double) * dof * dof);
  ORC_CATCH
}

// GP prior factor (unwhitened): err [2 dof], out_H [4][2 dof][dof] (may be NULL)
int orc_gp_prior(int dof, int lie, double delta_t, const double* x1, const double* v1, const double* x2,
                 const double* v2, double* out_err, double* out_H) {
  ORC_TRY
  Mat H[4];
  const Vec e = gp_prior_factor(dof, lie != 0, delta_t, x1, v1, x2, v2, out_H ? H : nullptr);
  std::memcpy(out_err, e.data(), sizeof(double) * 2 * dof);
  if (out_H)
    for (int k = 0; k < 4; k++) std::memcpy(out_H + (size_t)k * 2 * dof * dof, H[k].a.data(), sizeof(double) * 2 * dof * dof);
  ORC_CATCH
}

// Lambda, Psi [2dof][2dof] (calcLambda / calcPsi)
int orc_gp_lambda_psi(int dof, const double* Qc, double delta_t, double tau, double* out_Lambda, double* out_Psi) {
  ORC_TRY
  Mat Qm = Mat::Identity(dof);
  if (Qc) for (int i = 0; i < dof * dof; i++) Qm.a[i] = Qc[i];
  const Mat L = calcLambda(getQc(Qm), delta_t, tau), P = calcPsi(getQc(Qm), delta_t, tau);
  std::memcpy(out_Lambda, L.a.data(), sizeof(double) * L.a.size());
  std::memcpy(out_Psi, P.a.data(), sizeof(double) * P.a.size());
  ORC_CATCH
}

// Same outputs/layout as gpmp2b_linearize (block-tridiagonal extraction of the dense H), plus
// optional dense H [n][n] in variable order [x_0,v_0,x_1,v_1,...].
int orc_linearize(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const gpmp2b_setting* st, int64_t B,
                  const double* start_conf, const double* start_vel, const double* end_conf, const double* end_vel,
                  const double* traj, double* out_Hdiag, double* out_Hoff, double* out_g, double* out_error,
                  double* out_dense_H) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  const int D = st->dof, N = st->total_step + 1, b = 2 * D, n = N * b;
  for (int64_t p = 0; p < B; p++) {
    Problem P(rb, f, *st, start_conf + p * D, start_vel + p * D, end_conf + p * D, end_vel + p * D);
    if (st->fix_enabled) { P.fix_conf = st->fix_conf + p * D; P.fix_vel = st->fix_vel + p * D; }
    Vec t(traj + p * 2 * N * D, traj + (p + 1) * 2 * N * D);
    std::vector<LinFactor> lin;
    P.linearize(t, lin);
    Mat H;
    Vec g;
    normal_equations(lin, 2 * N, D, H, g);
    if (out_dense_H) std::memcpy(out_dense_H + (size_t)p * n * n, H.a.data(), sizeof(double) * n * n);
    for (int i = 0; i < N; i++) {
      if (out_Hdiag)
        for (int r = 0; r < b; r++)
          for (int c = 0; c < b; c++) out_Hdiag[(((size_t)p * N + i) * b + r) * b + c] = H(i * b + r, i * b + c);
      if (out_Hoff && i < N - 1)
        for (int r = 0; r < b; r++)
          for (int c = 0; c < b; c++)
            out_Hoff[(((size_t)p * (N - 1) + i) * b + r) * b + c] = H(i * b + r, (i + 1) * b + c);
      if (out_g)
        for (int r = 0; r < b; r++) out_g[((size_t)p * N + i) * b + r] = g[i * b + r];
    }
    if (out_error) out_error[p] = P.error(t);
  }
  ORC_CATCH
}

// out_err [B][C][S], out_centers [B][C][S][3] (may be NULL): same layout as gpmp2b_obstacle_errors
int orc_obstacle_errors(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const gpmp2b_setting* st, int64_t B,
                        const double* traj, double* out_err, double* out_centers) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  const int D = st->dof, N = st->total_step + 1, K = st->obs_check_inter, S = rb.nr_spheres();
  const int C = N + (N - 1) * K;
  const double zeros[GPMP2B_MAX_DOF] = {0};
  for (int64_t p = 0; p < B; p++) {
    Problem P(rb, f, *st, zeros, zeros, zeros, zeros);
    Vec t(traj + p * 2 * N * D, traj + (p + 1) * 2 * N * D);
    int c = 0;
    for (int i = 0; i < N; i++) {
      for (int j = 0; j <= (i < N - 1 ? K : 0); j++, c++) {
        Vec conf(P.X(t, i), P.X(t, i) + D);
        if (j > 0) conf = P.interps[j - 1].interpolatePose(P.X(t, i), P.V(t, i), P.X(t, i + 1), P.V(t, i + 1), nullptr);
        const Vec e = obstacle_factor(rb, f, conf.data(), st->epsilon, nullptr);
        std::memcpy(out_err + ((size_t)p * C + c) * S, e.data(), sizeof(double) * S);
        if (out_centers) {
          Vec ctr;
          sphere_centers(rb, conf.data(), ctr, nullptr);
          std::memcpy(out_centers + ((size_t)p * C + c) * S * 3, ctr.data(), sizeof(double) * S * 3);
        }
      }
    }
  }
  ORC_CATCH
}

// Batched optimize on `nthreads` host threads (one problem per thread at a time -- the reference's
// "one BatchTrajOptimize call per core").  Same per-problem outputs as gpmp2b_batch_optimize, plus
// optional counters out_counts [B][3] = (linearizations, solves, error evaluations).
int orc_batch_optimize(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const gpmp2b_setting* st, int64_t B,
                       const double* start_conf, const double* start_vel, const double* end_conf,
                       const double* end_vel, const double* init_traj, double* out_traj, double* out_error,
                       double* out_coll_cost, int32_t* out_iters, int32_t* out_status, int64_t* out_counts,
                       int nthreads, int dense) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  const int D = st->dof, N = st->total_step + 1;
  const size_t TL = (size_t)2 * N * D;
  std::atomic<int64_t> next(0);
  std::atomic<int> failed(0);
  auto work = [&]() {
    for (;;) {
      const int64_t p = next.fetch_add(1);
      if (p >= B) break;
      try {
        Problem P(rb, f, *st, start_conf + p * D, start_vel + p * D, end_conf + p * D, end_vel + p * D);
        if (st->fix_enabled) { P.fix_conf = st->fix_conf + p * D; P.fix_vel = st->fix_vel + p * D; }
        Vec init(init_traj + p * TL, init_traj + (p + 1) * TL);
        OptResult r = optimize(P, init, dense != 0);
        std::memcpy(out_traj + p * TL, r.traj.data(), sizeof(double) * TL);
        if (out_error) out_error[p] = r.error;
        if (out_coll_cost) out_coll_cost[p] = collision_cost(rb, f, D, N, r.traj.data());
        if (out_iters) out_iters[p] = r.iters;
        if (out_status) out_status[p] = r.status;
        if (out_counts) { out_counts[3 * p] = r.n_lin; out_counts[3 * p + 1] = r.n_solve; out_counts[3 * p + 2] = r.n_err; }
      } catch (const std::exception& e) {
        std::fprintf(stderr, "[oracle] problem %ld: %s\n", (long)p, e.what());
        failed = 1;
      }
    }
  };
  if (nthreads <= 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; i++) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  if (failed) return -1;
  ORC_CATCH
}

int orc_collision_cost(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const gpmp2b_setting* st, int64_t B,
                       const double* traj, double* out_cost) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  const int D = st->dof, N = st->total_step + 1;
  for (int64_t p = 0; p < B; p++) out_cost[p] = collision_cost(rb, f, D, N, traj + p * 2 * N * D);
  ORC_CATCH
}

// graph error of a trajectory (NonlinearFactorGraph::error)
int orc_graph_error(const gpmp2b_robot_desc* rd, const gpmp2b_sdf_desc* sd, const gpmp2b_setting* st, int64_t B,
                    const double* start_conf, const double* start_vel, const double* end_conf, const double* end_vel,
                    const double* traj, double* out_error) {
  ORC_TRY
  Robot rb(*rd);
  Sdf f(*sd);
  const int D = st->dof, N = st->total_step + 1;
  for (int64_t p = 0; p < B; p++) {
    Problem P(rb, f, *st, start_conf + p * D, start_vel + p * D, end_conf + p * D, end_vel + p * D);
    if (st->fix_enabled) { P.fix_conf = st->fix_conf + p * D; P.fix_vel = st->fix_vel + p * D; }
    Vec t(traj + p * 2 * N * D, traj + (p + 1) * 2 * N * D);
    out_error[p] = P.error(t);
  }
  ORC_CATCH
}

// Pose2 helpers exposed for the golden tests of the mobile-arm path
// initArmTrajStraightLine (gpmp2/planner/TrajUtils.cpp:23-48) / initPose2VectorTrajStraightLine (:51-73);
// interpolate<Pose2>(a, b, t) = a * Expmap(t * Logmap(between(a, b)))  [GTSAM Lie.h, recalled]
int orc_init_straight_line(int lie, int dof, int total_step, int64_t B, const double* start, const double* end, double* out) {
  ORC_TRY
  const int N = total_step + 1;
  for (int64_t p = 0; p < B; p++) {
    const double *s = start + p * dof, *e = end + p * dof;
    double* t = out + p * 2 * N * dof;
    for (int i = 0; i <= total_step; i++) {
      double* x = t + (size_t)i * dof;
      double* v = t + (size_t)(N + i) * dof;
      const double ratio = static_cast<double>(i) / static_cast<double>(total_step);
      if (!lie) {
        for (int d = 0; d < dof; d++)
          x[d] = (i == 0) ? s[d] : (i == total_step) ? e[d] : ratio * e[d] + (1.0 - ratio) * s[d];
      } else {
        const Pose2 a{s[0], s[1], s[2]}, b{e[0], e[1], e[2]};
        double lg[3];
        p2_logmap(p2_compose(p2_inverse(a), b), lg);
        for (int k = 0; k < 3; k++) lg[k] *= ratio;
        const Pose2 q = p2_compose(a, p2_expmap(lg));
        x[0] = q.x; x[1] = q.y; x[2] = q.th;
        for (int d = 3; d < dof; d++) x[d] = (1.0 - ratio) * s[d] + ratio * e[d];
      }
      for (int d = 0; d < dof; d++) v[d] = (e[d] - s[d]) / static_cast<double>(total_step);
    }
  }
  ORC_CATCH
}

// interpolateArmTraj (TrajUtils.cpp:158-196) / interpolatePose2MobileArmTraj (:199-237)
int orc_interpolate_traj(int lie, int dof, int total_step, double delta_t, const double* Qc, int inter_step, int start_index,
                         int end_index, int64_t B, const double* traj, double* out) {
  ORC_TRY
  Mat Qm = Mat::Identity(dof);
  if (Qc) for (int i = 0; i < dof * dof; i++) Qm.a[i] = Qc[i];
  const int N = total_step + 1, Nout = (end_index - start_index) * (inter_step + 1) + 1;
  const double inter_dt = delta_t / static_cast<double>(inter_step + 1);
  std::vector<Interp> gps;
  for (int j = 1; j <= inter_step; j++) gps.emplace_back(Qm, delta_t, static_cast<double>(j) * inter_dt, lie != 0);
  for (int64_t p = 0; p < B; p++) {
    const double* t = traj + p * 2 * N * dof;
    double* o = out + p * 2 * Nout * dof;
    int ri = 0;
    auto put = [&](const double* x, const double* v) {
      std::memcpy(o + (size_t)ri * dof, x, sizeof(double) * dof);
      std::memcpy(o + (size_t)(Nout + ri) * dof, v, sizeof(double) * dof);
      ri++;
    };
    for (int i = start_index; i < end_index; i++) {
      const double *x1 = t + (size_t)i * dof, *v1 = t + (size_t)(N + i) * dof, *x2 = x1 + dof, *v2 = v1 + dof;
      put(x1, v1);
      for (int j = 1; j <= inter_step; j++) {
        const Vec c = gps[j - 1].interpolatePose(x1, v1, x2, v2, nullptr);
        const Vec w = gps[j - 1].interpolateVelocity(x1, v1, x2, v2);
        put(c.data(), w.data());
      }
    }
    put(t + (size_t)end_index * dof, t + (size_t)(N + end_index) * dof);
  }
  ORC_CATCH
}

// Pose2 group operations [GTSAM Pose2, recalled]; pinned by gpmp2/geometry/tests/testPose2Vector.cpp:61-113
int orc_pose2_compose(const double* a, const double* b, double* out) {
  const Pose2 r = p2_compose(Pose2{a[0], a[1], a[2]}, Pose2{b[0], b[1], b[2]});
  out[0] = r.x; out[1] = r.y; out[2] = r.th; return 0;
}
int orc_pose2_between(const double* a, const double* b, double* out) {
  const Pose2 r = p2_compose(p2_inverse(Pose2{a[0], a[1], a[2]}), Pose2{b[0], b[1], b[2]});
  out[0] = r.x; out[1] = r.y; out[2] = r.th; return 0;
}
int orc_pose2_inverse(const double* a, double* out) {
  const Pose2 r = p2_inverse(Pose2{a[0], a[1], a[2]});
  out[0] = r.x; out[1] = r.y; out[2] = r.th; return 0;
}
int orc_pose2_expmap(const double* v, double* out) { const Pose2 p = p2_expmap(v); out[0] = p.x; out[1] = p.y; out[2] = p.th; return 0; }
int orc_pose2_logmap(const double* p, double* out) { p2_logmap(Pose2{p[0], p[1], p[2]}, out); return 0; }

}  // extern "C"
