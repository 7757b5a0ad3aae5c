// gpmp2.hpp -- header-only C++ facade with the reference's names and signatures over the C ABI (gpmp2b.h).
//
// A user of ori-drs/gpmp2's batch planner switches by including this header and linking libgpmp2b.so:
//   gpmp2::Arm, BodySphere, ArmModel            gpmp2/kinematics/Arm.h:47-55, RobotModel.h:20-27,56, ArmModel.h:19
//   gpmp2::Pose2MobileArm, Pose2MobileArmModel  gpmp2/kinematics/Pose2MobileArm.h:41, Pose2MobileArmModel.h:19
//   gpmp2::PlanarSDF, SignedDistanceField       gpmp2/obstacle/PlanarSDF.h:45-47, SignedDistanceField.h:58-79
//   gpmp2::TrajOptimizerSetting                 gpmp2/planner/TrajOptimizerSetting.h:17-100
//   gpmp2::BatchTrajOptimize2DArm/3DArm/Pose2MobileArm2D/Pose2MobileArm, CollisionCost*, optimize semantics
//                                               gpmp2/planner/BatchTrajOptimizer.h:43-185
//   gpmp2::initArmTrajStraightLine              gpmp2/planner/TrajUtils.h:28-30
// GTSAM types are replaced by light stand-ins (gtsam::Vector -> std::vector<double>, gtsam::Values -> a map
// from Symbol('x'|'v', i) to vectors); errors come back as the exceptions the reference throws
// (std::runtime_error).  Batched overloads (many problems per call) are what the hardware is for; the
// single-problem signatures are B = 1 calls of the same entry point.  There is no CPU fallback.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <vector>

#include "../gpmp2b.h"

namespace gpmp2 {

typedef std::vector<double> Vector;

struct Point3 {
  double x_, y_, z_;
  Point3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {}
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
};
struct Point2 {
  double x_, y_;
  Point2(double x = 0, double y = 0) : x_(x), y_(y) {}
  double x() const { return x_; }
  double y() const { return y_; }
};

/// rigid transform as a row-major 4x4 matrix (gtsam::Pose3::matrix())
struct Pose3 {
  std::array<double, 16> T;
  Pose3() : T{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}} {}
  Pose3(const std::array<double, 9>& R, const Point3& t) : Pose3() {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[3] = t.x(); T[7] = t.y(); T[11] = t.z();
  }
  static Pose3 Translation(const Point3& t) {
    Pose3 p;
    p.T[3] = t.x(); p.T[7] = t.y(); p.T[11] = t.z();
    return p;
  }
};

struct Pose2 {
  double x_, y_, theta_;
  Pose2(double x = 0, double y = 0, double theta = 0) : x_(x), y_(y), theta_(theta) {}
  double x() const { return x_; }
  double y() const { return y_; }
  double theta() const { return theta_; }
};

/// SE(2) x R^n state (gpmp2/geometry/Pose2Vector.h:26-53); flat wire form (x, y, theta, q...)
struct Pose2Vector {
  Pose2 pose_;
  Vector conf_;
  Pose2Vector() {}
  Pose2Vector(const Pose2& p, const Vector& c) : pose_(p), conf_(c) {}
  const Pose2& pose() const { return pose_; }
  const Vector& configuration() const { return conf_; }
  Vector flat() const {
    Vector v{pose_.x(), pose_.y(), pose_.theta()};
    v.insert(v.end(), conf_.begin(), conf_.end());
    return v;
  }
};

/// gtsam::Symbol(c, j)
struct Symbol {
  unsigned char c;
  std::uint64_t j;
  Symbol(unsigned char c_, std::uint64_t j_) : c(c_), j(j_) {}
  bool operator<(const Symbol& o) const { return c != o.c ? c < o.c : j < o.j; }
};

/// stand-in for gtsam::Values holding vectors (Pose2Vector values are stored flat)
class Values {
  std::map<Symbol, Vector> m_;
 public:
  void insert(const Symbol& k, const Vector& v) {
    if (!m_.insert(std::make_pair(k, v)).second) throw std::runtime_error("ValuesKeyAlreadyExists");
  }
  void insert(const Symbol& k, const Pose2Vector& v) { insert(k, v.flat()); }
  const Vector& at(const Symbol& k) const {
    auto it = m_.find(k);
    if (it == m_.end()) throw std::runtime_error("ValuesKeyDoesNotExist");
    return it->second;
  }
  bool exists(const Symbol& k) const { return m_.count(k) != 0; }
  size_t size() const { return m_.size(); }
};

/// gpmp2::insertPose2VectorInValues / atPose2VectorValues (gpmp2/utils/matlabUtils.cpp:14-22): the wrapper helpers that put a
/// Pose2Vector into / read it from the Values of a mobile-manipulator planner
inline void insertPose2VectorInValues(const Symbol& key, const Pose2Vector& p, Values& values) { values.insert(key, p); }
inline Pose2Vector atPose2VectorValues(const Symbol& key, const Values& values) {
  const Vector& v = values.at(key);
  if (v.size() < 3) throw std::runtime_error("atPose2VectorValues: the value under this key is not a Pose2Vector");
  return Pose2Vector(Pose2(v[0], v[1], v[2]), Vector(v.begin() + 3, v.end()));
}

// ------------------------------------------------------------------------------------------------
namespace detail {
inline void check(gpmp2b_ctx* ctx, int rc) {
  if (rc == GPMP2B_OK) return;
  const std::string msg = ctx ? gpmp2b_last_error(ctx) : "gpmp2b: no context";
  throw std::runtime_error(msg.empty() ? "gpmp2b error " + std::to_string(rc) : msg);
}
/// one process-wide context per device (the reference is single-threaded and synchronous)
inline gpmp2b_ctx* context(int device = 0) {
  struct Holder {
    gpmp2b_ctx* c = nullptr;
    ~Holder() { if (c) gpmp2b_destroy(c); }
  };
  static Holder h;
  if (!h.c) {
    const int rc = gpmp2b_create(device, &h.c);
    if (rc != GPMP2B_OK)
      throw std::runtime_error("gpmp2b_create failed (" + std::to_string(rc) + "): no usable CUDA device, and there is no CPU fallback");
  }
  return h.c;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------
class Arm {
  size_t dof_;
  Vector a_, alpha_, d_, theta_bias_;
  Pose3 base_pose_;
 public:
  Arm() : dof_(0) {}
  Arm(size_t dof, const Vector& a, const Vector& alpha, const Vector& d, const Pose3& base_pose = Pose3(),
      const Vector& theta_bias = Vector())
      : dof_(dof), a_(a), alpha_(alpha), d_(d), theta_bias_(theta_bias.empty() ? Vector(dof, 0.0) : theta_bias),
        base_pose_(base_pose) {
    if (a.size() != dof || alpha.size() != dof || d.size() != dof || theta_bias_.size() != dof)
      throw std::runtime_error("[Arm] ERROR: DH parameter dim does not fit dof.");
  }
  size_t dof() const { return dof_; }
  const Vector& a() const { return a_; }
  const Vector& d() const { return d_; }
  const Vector& alpha() const { return alpha_; }
  const Vector& theta_bias() const { return theta_bias_; }
  const Pose3& base_pose() const { return base_pose_; }
};

struct BodySphere {
  size_t link_id;
  double radius;
  Point3 center;
  BodySphere(size_t id, double r, const Point3& c) : link_id(id), radius(r), center(c) {}
};
typedef std::vector<BodySphere> BodySphereVector;

namespace detail {
/// device-resident robot, shared by copies of the model object
struct RobotHandle {
  gpmp2b_robot* h = nullptr;
  ~RobotHandle() { if (h) gpmp2b_robot_free(context(), h); }
};
inline std::shared_ptr<RobotHandle> upload_robot(int kind, const Arm& arm, const Pose3& base, const BodySphereVector& sph,
                                                 const Arm* arm2 = nullptr, bool reverse_linact = false,
                                                 const Pose3* base2 = nullptr, const Pose3* base3 = nullptr) {
  std::vector<int32_t> link(sph.size());
  Vector radius(sph.size()), center(3 * sph.size());
  for (size_t i = 0; i < sph.size(); i++) {
    link[i] = (int32_t)sph[i].link_id;
    radius[i] = sph[i].radius;
    center[3 * i] = sph[i].center.x(); center[3 * i + 1] = sph[i].center.y(); center[3 * i + 2] = sph[i].center.z();
  }
  gpmp2b_robot_desc d{};
  d.kind = kind; d.arm_dof = (int32_t)arm.dof(); d.n_spheres = (int32_t)sph.size();
  // two-arm robots: the DH tables of arm 1 followed by those of arm 2 (gpmp2b.h)
  Vector a = arm.a(), alpha = arm.alpha(), dd = arm.d(), bias = arm.theta_bias();
  if (arm2) {
    a.insert(a.end(), arm2->a().begin(), arm2->a().end());
    alpha.insert(alpha.end(), arm2->alpha().begin(), arm2->alpha().end());
    dd.insert(dd.end(), arm2->d().begin(), arm2->d().end());
    bias.insert(bias.end(), arm2->theta_bias().begin(), arm2->theta_bias().end());
    d.arm2_dof = (int32_t)arm2->dof();
  }
  d.reverse_linact = reverse_linact ? 1 : 0;
  d.a = a.data(); d.alpha = alpha.data(); d.d = dd.data(); d.theta_bias = bias.data();
  for (int i = 0; i < 16; i++) d.base_pose[i] = base.T[i];
  if (base2) for (int i = 0; i < 16; i++) d.base_pose2[i] = base2->T[i];
  if (base3) for (int i = 0; i < 16; i++) d.base_pose3[i] = base3->T[i];
  d.sphere_link = link.data(); d.sphere_radius = radius.data(); d.sphere_center = center.data();
  auto h = std::make_shared<RobotHandle>();
  check(context(), gpmp2b_robot_upload(context(), &d, &h->h));
  return h;
}
}  // namespace detail

/// RobotModel<Arm>
struct Matrix;

class ArmModel {
  Arm arm_;
  BodySphereVector spheres_;
  std::shared_ptr<detail::RobotHandle> dev_;
 public:
  typedef Vector Pose;
  typedef Vector Velocity;
  ArmModel(const Arm& arm, const BodySphereVector& spheres)
      : arm_(arm), spheres_(spheres), dev_(detail::upload_robot(GPMP2B_ROBOT_ARM, arm, arm.base_pose(), spheres)) {}
  const Arm& fk_model() const { return arm_; }
  size_t dof() const { return arm_.dof(); }
  size_t nr_body_spheres() const { return spheres_.size(); }
  size_t sphere_link_id(size_t i) const { return spheres_[i].link_id; }
  double sphere_radius(size_t i) const { return spheres_[i].radius; }
  /// RobotModel::sphereCentersMat (gpmp2/kinematics/RobotModel-inl.h:71-82): 3 x S sphere centres in the world frame (device)
  Matrix sphereCentersMat(const Vector& conf) const;
  const gpmp2b_robot* device() const { return dev_->h; }
};

class Pose2MobileArm {
  Pose3 base_T_arm_;
  Arm arm_;
 public:
  explicit Pose2MobileArm(const Arm& arm, const Pose3& base_T_arm = Pose3()) : base_T_arm_(base_T_arm), arm_(arm) {}
  size_t dof() const { return arm_.dof() + 3; }
  size_t nr_links() const { return arm_.dof() + 1; }
  const Pose3& base_T_arm() const { return base_T_arm_; }
  const Arm& arm() const { return arm_; }
};

/// RobotModel<Pose2MobileArm>
class Pose2MobileArmModel {
  Pose2MobileArm marm_;
  BodySphereVector spheres_;
  std::shared_ptr<detail::RobotHandle> dev_;
 public:
  typedef Pose2Vector Pose;
  typedef Vector Velocity;
  Pose2MobileArmModel(const Pose2MobileArm& marm, const BodySphereVector& spheres)
      : marm_(marm), spheres_(spheres),
        dev_(detail::upload_robot(GPMP2B_ROBOT_POSE2_MOBILE_ARM, marm.arm(), marm.base_T_arm(), spheres)) {}
  const Pose2MobileArm& fk_model() const { return marm_; }
  size_t dof() const { return marm_.dof(); }
  size_t nr_body_spheres() const { return spheres_.size(); }
  Matrix sphereCentersMat(const Pose2Vector& conf) const;
  const gpmp2b_robot* device() const { return dev_->h; }
};

/// gpmp2::Pose2Mobile2Arms (gpmp2/kinematics/Pose2Mobile2Arms.h:23-66): vehicle + two arms
class Pose2Mobile2Arms {
  Pose3 base_T_arm1_, base_T_arm2_;
  Arm arm1_, arm2_;
 public:
  Pose2Mobile2Arms(const Arm& arm1, const Arm& arm2, const Pose3& base_T_arm1 = Pose3(), const Pose3& base_T_arm2 = Pose3())
      : base_T_arm1_(base_T_arm1), base_T_arm2_(base_T_arm2), arm1_(arm1), arm2_(arm2) {}
  size_t dof() const { return arm1_.dof() + arm2_.dof() + 3; }
  size_t nr_links() const { return arm1_.dof() + arm2_.dof() + 1; }
  const Pose3& base_T_arm1() const { return base_T_arm1_; }
  const Pose3& base_T_arm2() const { return base_T_arm2_; }
  const Arm& arm1() const { return arm1_; }
  const Arm& arm2() const { return arm2_; }
};
/// gpmp2::Pose2MobileVetLinArm (gpmp2/kinematics/Pose2MobileVetLinArm.h:24-68): vehicle + vertical linear actuator + arm
class Pose2MobileVetLinArm {
  Pose3 base_T_torso_, torso_T_arm_;
  bool reverse_linact_;
  Arm arm_;
 public:
  explicit Pose2MobileVetLinArm(const Arm& arm, const Pose3& base_T_torso = Pose3(), const Pose3& torso_T_arm = Pose3(), bool reverse_linact = false)
      : base_T_torso_(base_T_torso), torso_T_arm_(torso_T_arm), reverse_linact_(reverse_linact), arm_(arm) {}
  size_t dof() const { return arm_.dof() + 4; }
  size_t nr_links() const { return arm_.dof() + 2; }
  const Pose3& base_T_torso() const { return base_T_torso_; }
  const Pose3& torso_T_arm() const { return torso_T_arm_; }
  bool reverse_linact() const { return reverse_linact_; }
  const Arm& arm() const { return arm_; }
};
/// gpmp2::Pose2MobileVetLin2Arms (gpmp2/kinematics/Pose2MobileVetLin2Arms.h:24-75): vehicle + linear actuator + two arms
class Pose2MobileVetLin2Arms {
  Pose3 base_T_torso_, torso_T_arm1_, torso_T_arm2_;
  bool reverse_linact_;
  Arm arm1_, arm2_;
 public:
  Pose2MobileVetLin2Arms(const Arm& arm1, const Arm& arm2, const Pose3& base_T_torso = Pose3(), const Pose3& torso_T_arm1 = Pose3(),
                         const Pose3& torso_T_arm2 = Pose3(), bool reverse_linact = false)
      : base_T_torso_(base_T_torso), torso_T_arm1_(torso_T_arm1), torso_T_arm2_(torso_T_arm2), reverse_linact_(reverse_linact),
        arm1_(arm1), arm2_(arm2) {}
  size_t dof() const { return arm1_.dof() + arm2_.dof() + 4; }
  size_t nr_links() const { return arm1_.dof() + arm2_.dof() + 2; }
  const Pose3& base_T_torso() const { return base_T_torso_; }
  const Pose3& torso_T_arm1() const { return torso_T_arm1_; }
  const Pose3& torso_T_arm2() const { return torso_T_arm2_; }
  bool reverse_linact() const { return reverse_linact_; }
  const Arm& arm1() const { return arm1_; }
  const Arm& arm2() const { return arm2_; }
};

/// RobotModel<FK> of the three robots above (Pose2Mobile2ArmsModel.h, Pose2MobileVetLinArmModel.h, Pose2MobileVetLin2ArmsModel.h)
template <class FK>
class MobileModelT {
  FK fk_;
  BodySphereVector spheres_;
  std::shared_ptr<detail::RobotHandle> dev_;
 public:
  typedef Pose2Vector Pose;
  typedef Vector Velocity;
  MobileModelT(const FK& fk, const BodySphereVector& spheres, std::shared_ptr<detail::RobotHandle> dev) : fk_(fk), spheres_(spheres), dev_(dev) {}
  const FK& fk_model() const { return fk_; }
  size_t dof() const { return fk_.dof(); }
  size_t nr_body_spheres() const { return spheres_.size(); }
  Matrix sphereCentersMat(const Pose2Vector& conf) const;
  const gpmp2b_robot* device() const { return dev_->h; }
};
struct Pose2Mobile2ArmsModel : MobileModelT<Pose2Mobile2Arms> {
  Pose2Mobile2ArmsModel(const Pose2Mobile2Arms& m, const BodySphereVector& sph)
      : MobileModelT(m, sph, detail::upload_robot(GPMP2B_ROBOT_POSE2_MOBILE_2ARMS, m.arm1(), m.base_T_arm1(), sph, &m.arm2(), false, &m.base_T_arm2())) {}
};
struct Pose2MobileVetLinArmModel : MobileModelT<Pose2MobileVetLinArm> {
  Pose2MobileVetLinArmModel(const Pose2MobileVetLinArm& m, const BodySphereVector& sph)
      : MobileModelT(m, sph, detail::upload_robot(GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_ARM, m.arm(), m.base_T_torso(), sph, nullptr, m.reverse_linact(), &m.torso_T_arm())) {}
};
struct Pose2MobileVetLin2ArmsModel : MobileModelT<Pose2MobileVetLin2Arms> {
  Pose2MobileVetLin2ArmsModel(const Pose2MobileVetLin2Arms& m, const BodySphereVector& sph)
      : MobileModelT(m, sph, detail::upload_robot(GPMP2B_ROBOT_POSE2_MOBILE_VETLIN_2ARMS, m.arm1(), m.base_T_torso(), sph, &m.arm2(), m.reverse_linact(),
                                                  &m.torso_T_arm1(), &m.torso_T_arm2())) {}
};

// ------------------------------------------------------------------------------------------------
/// row-major dense matrix stand-in for gtsam::Matrix: data(r, c) = field at y index r, x index c
struct Matrix {
  size_t rows_, cols_;
  std::vector<double> a;
  Matrix(size_t r = 0, size_t c = 0) : rows_(r), cols_(c), a(r * c, 0.0) {}
  double& operator()(size_t r, size_t c) { return a[r * cols_ + c]; }
  double operator()(size_t r, size_t c) const { return a[r * cols_ + c]; }
  size_t rows() const { return rows_; }
  size_t cols() const { return cols_; }
};

namespace detail {
struct SdfHandle {
  gpmp2b_sdf* h = nullptr;
  ~SdfHandle() { if (h) gpmp2b_sdf_free(context(), h); }
};
}  // namespace detail

class PlanarSDF {
  Point2 origin_;
  size_t rows_, cols_;
  double cell_size_;
  std::shared_ptr<detail::SdfHandle> dev_;
 public:
  PlanarSDF(const Point2& origin, double cell_size, const Matrix& data)
      : origin_(origin), rows_(data.rows()), cols_(data.cols()), cell_size_(cell_size), dev_(std::make_shared<detail::SdfHandle>()) {
    std::vector<double> wire(rows_ * cols_);   // [col][row]
    for (size_t r = 0; r < rows_; r++)
      for (size_t c = 0; c < cols_; c++) wire[c * rows_ + r] = data(r, c);
    gpmp2b_sdf_desc d{};
    d.ndim = 2; d.rows = (int32_t)rows_; d.cols = (int32_t)cols_; d.nz = 1;
    d.origin[0] = origin.x(); d.origin[1] = origin.y(); d.origin[2] = 0.0;
    d.cell_size = cell_size; d.data = wire.data();
    detail::check(detail::context(), gpmp2b_sdf_upload(detail::context(), &d, &dev_->h));
  }
  size_t x_count() const { return cols_; }
  size_t y_count() const { return rows_; }
  double cell_size() const { return cell_size_; }
  const Point2& origin() const { return origin_; }
  const gpmp2b_sdf* device() const { return dev_->h; }
};

class SignedDistanceField {
  Point3 origin_;
  size_t rows_, cols_, z_;
  double cell_size_;
  std::vector<double> wire_;   // [z][col][row]
  mutable std::shared_ptr<detail::SdfHandle> dev_;
 public:
  SignedDistanceField(const Point3& origin, double cell_size, const std::vector<Matrix>& data)
      : origin_(origin), rows_(data.at(0).rows()), cols_(data.at(0).cols()), z_(data.size()), cell_size_(cell_size),
        wire_(rows_ * cols_ * z_) {
    for (size_t z = 0; z < z_; z++) initFieldData(z, data[z]);
  }
  SignedDistanceField(const Point3& origin, double cell_size, size_t field_rows, size_t field_cols, size_t field_z)
      : origin_(origin), rows_(field_rows), cols_(field_cols), z_(field_z), cell_size_(cell_size), wire_(rows_ * cols_ * z_, 0.0) {}
  void initFieldData(size_t z_idx, const Matrix& field_layer) {
    if (z_idx >= z_) throw std::runtime_error("[SignedDistanceField] matrix layer out of index");
    for (size_t r = 0; r < rows_; r++)
      for (size_t c = 0; c < cols_; c++) wire_[(z_idx * cols_ + c) * rows_ + r] = field_layer(r, c);
    dev_.reset();
  }
  size_t x_count() const { return cols_; }
  size_t y_count() const { return rows_; }
  size_t z_count() const { return z_; }
  double cell_size() const { return cell_size_; }
  const Point3& origin() const { return origin_; }
  /// raw field value (SignedDistanceField.h:170-172)
  double signed_distance(size_t r, size_t c, size_t z) const { return wire_[(z * cols_ + c) * rows_ + r]; }

  /// SignedDistanceField::saveSDF / loadSDF (gpmp2/obstacle/SignedDistanceField.cpp:14-50): Boost.Serialization archive
  /// picked by the extension -- ".bin" binary_oarchive, anything else text_oarchive -- of the members in the order of
  /// SignedDistanceField.h:201-208.  The archive framing is pinned to a real Boost 1.78 runtime (tests/golden/sdf_boost178_*,
  /// written through oracle/boost_probe); the member lists -- GTSAM's Eigen serialization (rows, cols, column-major
  /// coefficients), gtsam::Point3 -- are restated: no GTSAM here (library version >= 7, little-endian LP64).  Every class costs (tracking level, version) once: two
  /// integers in text, 1 + 4 bytes in binary.  gtsam::Point3 has had three layouts -- A: class Point3 : Vector3
  /// (GTSAM 4.0), B: typedef of Vector3, C: x_, y_, z_ members (GTSAM 3) -- the reader takes whichever accounts for the
  /// whole file, the writer emits A.  ".xml": the reference wraps BOOST_SERIALIZATION_NVP(*this), a tag name Boost's
  /// xml archive rejects, so it cannot write one either; it throws here too.  Same format as gpmp2_b200/boost_archive.py.
  void saveSDF(const std::string& filename) const {
    const std::string ext = filename.substr(filename.find_last_of(".") + 1);
    if (ext == "xml") throw std::runtime_error("[saveSDF] .xml: tag name '*this' is not a valid XML name (the reference throws too)");
    const bool bin = ext == "bin";
    std::ofstream f(filename.c_str(), bin ? std::ios::binary : std::ios::out);
    if (!f.good()) throw std::runtime_error("[saveSDF] cannot open '" + filename + "'");
    const uint64_t n = rows_ * cols_;
    if (bin) {
      auto put = [&f](const void* p, size_t k) { f.write(static_cast<const char*>(p), (std::streamsize)k); };
      auto u64 = [&put](uint64_t v) { put(&v, 8); };
      auto cls = [&put]() { const unsigned char c[5] = {0, 0, 0, 0, 0}; put(c, 5); };
      const char sig[] = "serialization::archive";
      const uint16_t ver = 17;
      const unsigned char sizes[4] = {4, 8, 4, 8};
      const int32_t one = 1;
      const uint32_t item_version = 0;
      u64(22); put(sig, 22); put(&ver, 2); put(sizes, 4); put(&one, 4);
      cls(); cls(); cls(); u64(3); u64(1);
      const double o[3] = {origin_.x(), origin_.y(), origin_.z()};
      put(o, 24); u64(rows_); u64(cols_); u64(z_); put(&cell_size_, 8);
      cls(); u64(z_); put(&item_version, 4);
      for (size_t z = 0; z < z_; z++) {
        if (z == 0) cls();
        u64(rows_); u64(cols_); put(wire_.data() + z * n, 8 * n);
      }
    } else {
      char b[40];
      auto dbl = [&f, &b](double v) { std::snprintf(b, sizeof b, " %.17e", v); f << b; };
      f << "22 serialization::archive 17 0 0 0 0 0 0 3 1";
      dbl(origin_.x()); dbl(origin_.y()); dbl(origin_.z());
      f << ' ' << rows_ << ' ' << cols_ << ' ' << z_;
      dbl(cell_size_);
      f << " 0 0 " << z_ << " 0";
      for (size_t z = 0; z < z_; z++) {
        f << (z == 0 ? " 0 0 " : " ") << rows_ << ' ' << cols_;
        for (uint64_t i = 0; i < n; i++) dbl(wire_[z * n + i]);
      }
      f << "\n";
    }
    if (!f.good()) throw std::runtime_error("[saveSDF] write to '" + filename + "' failed");
  }

  void loadSDF(const std::string& filename) {
    const std::string ext = filename.substr(filename.find_last_of(".") + 1);
    if (ext == "xml") throw std::runtime_error("[loadSDF] .xml: tag name '*this' is not a valid XML name (the reference cannot write one)");
    std::ifstream f(filename.c_str(), std::ios::binary);
    if (!f.good()) throw std::runtime_error("File '" + filename + "' does not exist!");
    const std::string raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    // numeric tokens of the body: text = whitespace-separated numbers; binary = walked with the same grammar
    std::vector<double> tok;      // text only
    size_t version = 0;
    const bool bin = ext == "bin";
    size_t body = 0;
    if (bin) {
      static const unsigned char head[8] = {22, 0, 0, 0, 0, 0, 0, 0};
      static const unsigned char sizes[8] = {4, 8, 4, 8, 1, 0, 0, 0};
      if (raw.size() < 40 || std::memcmp(raw.data(), head, 8) || raw.compare(8, 22, "serialization::archive"))
        throw std::runtime_error("[loadSDF] not a Boost binary archive (signature missing)");
      uint16_t v; std::memcpy(&v, raw.data() + 30, 2); version = v;
      if (version < 7) throw std::runtime_error("[loadSDF] binary archive of a Boost library version < 7 is not supported");
      if (std::memcmp(raw.data() + 32, sizes, 8))
        throw std::runtime_error("[loadSDF] binary archive written with other type sizes or byte order (not portable)");
      body = 40;
    } else {
      std::istringstream is(raw);
      size_t len; std::string sig;
      if (!(is >> len >> sig >> version) || len != 22 || sig != "serialization::archive")
        throw std::runtime_error("[loadSDF] not a Boost text archive (signature missing)");
      double v;
      while (is >> v) tok.push_back(v);
      if (!is.eof()) throw std::runtime_error("[loadSDF] text archive holds a token that is not a number");
    }
    for (int layout = 0; layout < 3; layout++) {      // A, B, C
      size_t i = body;                                // byte offset (binary) or token index (text)
      bool ok = true;
      auto avail = [&](size_t k) { return i + k <= (bin ? raw.size() : tok.size()); };
      auto u = [&]() -> uint64_t {                    // size_t member
        if (!avail(bin ? 8 : 1)) { ok = false; return 0; }
        uint64_t v;
        if (bin) { std::memcpy(&v, raw.data() + i, 8); i += 8; }
        else {
          const double t = tok[i++];
          if (!(t >= 0.0 && t < 9.0e15) || t != std::floor(t)) { ok = false; return 0; }   // not a size
          v = (uint64_t)t;
        }
        return v;
      };
      auto d = [&]() -> double {
        if (!avail(bin ? 8 : 1)) { ok = false; return 0; }
        double v;
        if (bin) { std::memcpy(&v, raw.data() + i, 8); i += 8; } else { v = tok[i++]; }
        return v;
      };
      auto cls = [&]() {                              // tracking level 0, class version 0
        if (bin) { if (!avail(5) || raw.compare(i, 5, std::string(5, '\0'))) ok = false; i += 5; }
        else { if (!avail(2) || tok[i] != 0 || tok[i + 1] != 0) ok = false; i += 2; }
      };
      cls(); cls();
      if (layout == 0) cls();
      if (layout <= 1 && (u() != 3 || u() != 1)) ok = false;
      const double ox = d(), oy = d(), oz = d();
      const uint64_t rows = u(), cols = u(), nz = u();
      const double cell = d();
      cls();
      if (u() != nz) ok = false;
      if (bin) { if (!avail(4)) ok = false; i += 4; } else if (version > 3) { i += 1; }       // item_version
      if (!ok) continue;
      const uint64_t n = rows * cols;
      if ((rows | cols | nz) >> 31 || n > raw.size() || nz > raw.size()) continue;   // more coefficients / layers than the file has bytes
      const uint64_t need = bin ? nz * (16 + 8 * n) + (nz ? 5 : 0) : nz * (2 + n) + (nz ? 2 : 0);
      if ((bin ? raw.size() : tok.size()) - i != need) continue;
      std::vector<double> wire(nz * n);
      for (uint64_t z = 0; z < nz && ok; z++) {
        if (z == 0) cls();
        if (u() != rows || u() != cols) ok = false;
        if (bin) { std::memcpy(wire.data() + z * n, raw.data() + i, 8 * n); i += 8 * n; }
        else { for (uint64_t k = 0; k < n; k++) wire[z * n + k] = tok[i++]; }
      }
      if (!ok) continue;
      origin_ = Point3(ox, oy, oz); rows_ = rows; cols_ = cols; z_ = nz; cell_size_ = cell; wire_.swap(wire);
      dev_.reset();
      return;
    }
    throw std::runtime_error("[loadSDF] archive does not hold a SignedDistanceField in any known layout");
  }

  const gpmp2b_sdf* device() const {
    if (!dev_) {
      dev_ = std::make_shared<detail::SdfHandle>();
      gpmp2b_sdf_desc d{};
      d.ndim = 3; d.rows = (int32_t)rows_; d.cols = (int32_t)cols_; d.nz = (int32_t)z_;
      d.origin[0] = origin_.x(); d.origin[1] = origin_.y(); d.origin[2] = origin_.z();
      d.cell_size = cell_size_; d.data = wire_.data();
      detail::check(detail::context(), gpmp2b_sdf_upload(detail::context(), &d, &dev_->h));
    }
    return dev_->h;
  }
};

/// signedDistanceField3D (matlab/+gpmp2/signedDistanceField3D.m:16-33) on the device: occupancy > 0.75 is an obstacle;
/// returns the field layers (rows x cols matrices, one per z) ready for SignedDistanceField(origin, cell_size, layers).
/// single_precision = true reproduces MATLAB's bwdist arithmetic, false stays in double.
inline std::vector<Matrix> signedDistanceField3D(const std::vector<Matrix>& ground_truth_map, double cell_size,
                                                 bool single_precision = true) {
  const size_t nz = ground_truth_map.size(), rows = ground_truth_map.at(0).rows(), cols = ground_truth_map.at(0).cols();
  std::vector<double> wire(rows * cols * nz), out(rows * cols * nz);   // [z][col][row]
  for (size_t z = 0; z < nz; z++)
    for (size_t r = 0; r < rows; r++)
      for (size_t c = 0; c < cols; c++) wire[(z * cols + c) * rows + r] = ground_truth_map[z](r, c);
  gpmp2b_sdf_desc d{};
  d.ndim = 3; d.rows = (int32_t)rows; d.cols = (int32_t)cols; d.nz = (int32_t)nz; d.cell_size = cell_size; d.data = wire.data();
  detail::check(detail::context(), gpmp2b_sdf_from_occupancy(detail::context(), &d, single_precision ? 1 : 0, nullptr, out.data()));
  std::vector<Matrix> layers(nz, Matrix(rows, cols));
  for (size_t z = 0; z < nz; z++)
    for (size_t r = 0; r < rows; r++)
      for (size_t c = 0; c < cols; c++) layers[z](r, c) = out[(z * cols + c) * rows + r];
  return layers;
}
/// signedDistanceField2D (matlab/+gpmp2/signedDistanceField2D.m) on the device
inline Matrix signedDistanceField2D(const Matrix& ground_truth_map, double cell_size, bool single_precision = true) {
  const size_t rows = ground_truth_map.rows(), cols = ground_truth_map.cols();
  std::vector<double> wire(rows * cols), out(rows * cols);   // [col][row]
  for (size_t r = 0; r < rows; r++)
    for (size_t c = 0; c < cols; c++) wire[c * rows + r] = ground_truth_map(r, c);
  gpmp2b_sdf_desc d{};
  d.ndim = 2; d.rows = (int32_t)rows; d.cols = (int32_t)cols; d.nz = 1; d.cell_size = cell_size; d.data = wire.data();
  detail::check(detail::context(), gpmp2b_sdf_from_occupancy(detail::context(), &d, single_precision ? 1 : 0, nullptr, out.data()));
  Matrix field(rows, cols);
  for (size_t r = 0; r < rows; r++)
    for (size_t c = 0; c < cols; c++) field(r, c) = out[c * rows + r];
  return field;
}

/// gpmp2::readSDFvolfile (gpmp2/utils/fileUtils.cpp:17-62): `<pre>.vol.head` = cols rows z, origin, cell size (text);
/// `<pre>.vol.data` = field values (text), x outermost, then y, z innermost.  false if a file cannot be opened.
inline bool readSDFvolfile(const std::string& filename_pre, SignedDistanceField& sdf) {
  std::ifstream head((filename_pre + ".vol.head").c_str());
  if (!head.is_open()) return false;
  size_t field_rows, field_cols, field_z;
  double ox, oy, oz, res;
  head >> field_cols >> field_rows >> field_z >> ox >> oy >> oz >> res;
  std::ifstream data((filename_pre + ".vol.data").c_str());
  if (!data.is_open()) return false;
  std::vector<Matrix> vmat(field_z, Matrix(field_rows, field_cols));
  for (size_t x = 0; x < field_cols; x++)
    for (size_t y = 0; y < field_rows; y++)
      for (size_t z = 0; z < field_z; z++) data >> vmat[z](y, x);
  sdf = SignedDistanceField(Point3(ox, oy, oz), res, vmat);
  return true;
}

// ------------------------------------------------------------------------------------------------
struct TrajOptimizerSetting {
  enum IterationType { GaussNewton, LM, Dogleg };
  enum VerbosityLevel { None, Error };
  size_t dof, total_step;
  double total_time, conf_prior_sigma, vel_prior_sigma;
  bool flag_pos_limit, flag_vel_limit;
  Vector joint_pos_limits_up, joint_pos_limits_down, vel_limits, pos_limit_thresh, vel_limit_thresh, pos_limit_sigma,
      vel_limit_sigma;
  double epsilon, cost_sigma;
  size_t obs_check_inter;
  Vector Qc;   // row-major dof x dof covariance
  IterationType opt_type;
  VerbosityLevel opt_verbosity;
  bool final_iter_no_increase;
  double rel_thresh;
  size_t max_iter;
  // optional workspace goal on x_T (see gpmp2b_setting in gpmp2b.h); not part of the reference's struct, whose
  // hand-built graphs add a GoalFactorArm / GaussianPriorWorkspacePositionArm (matlab/Arm3GoalReachExample.m:107)
  bool goal_enabled = false, goal_keep_end_prior = false;
  int goal_link = -1;            // -1 = last joint frame (GoalFactorArm)
  double goal_sigma = 1.0;
  double goal_pos[3] = {0.0, 0.0, 0.0};
  void set_workspace_goal(double x, double y, double z, double sigma, int link = -1, bool keep_end_conf_prior = false) {
    goal_enabled = true; goal_pos[0] = x; goal_pos[1] = y; goal_pos[2] = z; goal_sigma = sigma; goal_link = link;
    goal_keep_end_prior = keep_end_conf_prior;
  }

  // optional SelfCollisionArm factor on every support state (gpmp2/obstacle/SelfCollision.h:38-60):
  // rows of (sphere A id, sphere B id, epsilon, sigma), row-major
  Vector self_collision_data;
  void set_self_collision(const Vector& data_rows_of_4) { self_collision_data = data_rows_of_4; }
  // optional VehicleDynamicsFactorPose2Vector(x_i, v_i, sigma) on every support state of a Pose2MobileArm; 0 = off
  double vehicle_dynamics_sigma = 0.0;
  void set_vehicle_dynamics(double sigma) { vehicle_dynamics_sigma = sigma; }
  // optional GaussianPriorWorkspaceOrientationArm on support states first..last (last < 0 = total_step); R row-major
  bool orient_enabled = false;
  int orient_link = -1, orient_state_first = 0, orient_state_last = -1;
  double orient_sigma = 1.0;
  double orient_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  void set_workspace_orientation(const double (&R)[9], double sigma, int link = -1, int first_state = 0, int last_state = -1) {
    orient_enabled = true; orient_sigma = sigma; orient_link = link; orient_state_first = first_state; orient_state_last = last_state;
    for (int k = 0; k < 9; k++) orient_R[k] = R[k];
  }

  // goal as a full pose: GaussianPriorWorkspacePoseArm on x_T (gpmp2b_setting.goal_enabled = 2); R row-major
  double goal_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool goal_is_pose = false;
  void set_workspace_pose_goal(const double (&R)[9], double x, double y, double z, double sigma, int link = -1,
                               bool keep_end_conf_prior = false) {
    set_workspace_goal(x, y, z, sigma, link, keep_end_conf_prior);
    goal_is_pose = true;
    for (int k = 0; k < 9; k++) goal_R[k] = R[k];
  }

  // one workspace target PER PROBLEM of a batched call (B rows each; empty = the shared value above): the reference
  // attaches these factors per graph, so a batch of different queries carries different goals
  Vector goal_pos_batch, goal_R_batch, orient_R_batch;     // [B][3], [B][9] row-major, [B][9] row-major
  void set_workspace_goal_batch(const Vector& points_rows_of_3, const Vector& rotations_rows_of_9 = Vector()) {
    goal_pos_batch = points_rows_of_3; goal_R_batch = rotations_rows_of_9;
  }
  void set_workspace_orientation_batch(const Vector& rotations_rows_of_9) { orient_R_batch = rotations_rows_of_9; }

  // replanning re-solve (gpmp2b.h, fix_*): ISAM2TrajOptimizer::fixConfigAndVel (ISAM2TrajOptimizer-inl.h:160-168) for a batch
  // of replanning problems in lockstep -- PriorFactor(x_k, conf_fix[p]) + PriorFactor(v_k, vel_fix[p]) with the conf / vel
  // prior models; conf_fix, vel_fix hold B rows of dof values.  New goals = new end_conf / end_vel of the call
  // (changeGoalConfigAndVel), warm start = init values from the previous result (initValues).
  bool fix_enabled = false;
  size_t fix_state_index = 0;
  Vector fix_conf, fix_vel;
  void fixConfigAndVel(size_t state_idx, const Vector& conf_fix_rows, const Vector& vel_fix_rows) {
    fix_enabled = true; fix_state_index = state_idx; fix_conf = conf_fix_rows; fix_vel = vel_fix_rows;
  }
  void clearFixedState() { fix_enabled = false; fix_conf.clear(); fix_vel.clear(); }

  /// defaults: gpmp2/planner/TrajOptimizerSetting.cpp:44-68
  explicit TrajOptimizerSetting(size_t system_dof)
      : dof(system_dof), total_step(10), total_time(1.0), conf_prior_sigma(0.0001), vel_prior_sigma(0.0001),
        flag_pos_limit(false), flag_vel_limit(false), joint_pos_limits_up(system_dof, 1e6),
        joint_pos_limits_down(system_dof, -1e6), vel_limits(system_dof, 1e6), pos_limit_thresh(system_dof, 0.001),
        vel_limit_thresh(system_dof, 0.001), pos_limit_sigma(system_dof, 0.001), vel_limit_sigma(system_dof, 0.001),
        epsilon(0.2), cost_sigma(0.1), obs_check_inter(5), Qc(system_dof * system_dof, 0.0), opt_type(Dogleg),
        opt_verbosity(None), final_iter_no_increase(true), rel_thresh(1e-2), max_iter(50) {
    for (size_t i = 0; i < dof; i++) Qc[i * dof + i] = 1.0;
  }
  void set_total_step(size_t step) { total_step = step; }
  void set_total_time(double time) { total_time = time; }
  void set_conf_prior_model(double sigma) { conf_prior_sigma = sigma; }
  void set_vel_prior_model(double sigma) { vel_prior_sigma = sigma; }
  void set_flag_pos_limit(bool flag) { flag_pos_limit = flag; }
  void set_flag_vel_limit(bool flag) { flag_vel_limit = flag; }
  void set_joint_pos_limits_up(const Vector& v) { joint_pos_limits_up = v; }
  void set_joint_pos_limits_down(const Vector& v) { joint_pos_limits_down = v; }
  void set_vel_limits(const Vector& v) { vel_limits = v; }
  void set_pos_limit_thresh(const Vector& v) { pos_limit_thresh = v; }
  void set_vel_limit_thresh(const Vector& v) { vel_limit_thresh = v; }
  void set_pos_limit_model(const Vector& v) { pos_limit_sigma = v; }
  void set_vel_limit_model(const Vector& v) { vel_limit_sigma = v; }
  void set_epsilon(double eps) { epsilon = eps; }
  void set_cost_sigma(double sigma) { cost_sigma = sigma; }
  void set_obs_check_inter(size_t inter) { obs_check_inter = inter; }
  void set_Qc_model(const Vector& Qc_rowmajor) { Qc = Qc_rowmajor; }
  void setGaussNewton() { opt_type = GaussNewton; }
  void setLM() { opt_type = LM; }
  void setDogleg() { opt_type = Dogleg; }
  void set_rel_thresh(double thresh) { rel_thresh = thresh; }
  void set_max_iter(size_t iter) { max_iter = iter; }
  void setVerbosityNone() { opt_verbosity = None; }
  void setVerbosityError() { opt_verbosity = Error; }
  void setOptimizationNoIncrase(bool flag) { final_iter_no_increase = flag; }

  gpmp2b_setting pack() const {
    auto need = [&](const Vector& v, const char* what) {
      if (v.size() != dof) throw std::runtime_error(std::string("[TrajOptimizerSetting] ERROR: ") + what + " dim does not fit.");
    };
    need(joint_pos_limits_up, "joint_pos_limits_up"); need(joint_pos_limits_down, "joint_pos_limits_down");
    need(vel_limits, "vel_limits"); need(pos_limit_thresh, "pos_limit_thresh"); need(vel_limit_thresh, "vel_limit_thresh");
    need(pos_limit_sigma, "pos_limit_model"); need(vel_limit_sigma, "vel_limit_model");
    if (Qc.size() != dof * dof) throw std::runtime_error("[TrajOptimizerSetting] ERROR: Qc dim does not fit.");
    gpmp2b_setting s{};
    s.dof = (int32_t)dof; s.total_step = (int32_t)total_step; s.total_time = total_time;
    s.conf_prior_sigma = conf_prior_sigma; s.vel_prior_sigma = vel_prior_sigma;
    s.flag_pos_limit = flag_pos_limit; s.flag_vel_limit = flag_vel_limit;
    s.joint_pos_limits_up = joint_pos_limits_up.data(); s.joint_pos_limits_down = joint_pos_limits_down.data();
    s.vel_limits = vel_limits.data(); s.pos_limit_thresh = pos_limit_thresh.data(); s.vel_limit_thresh = vel_limit_thresh.data();
    s.pos_limit_sigma = pos_limit_sigma.data(); s.vel_limit_sigma = vel_limit_sigma.data();
    s.epsilon = epsilon; s.cost_sigma = cost_sigma; s.obs_check_inter = (int32_t)obs_check_inter;
    s.opt_type = opt_type == GaussNewton ? GPMP2B_OPT_GAUSS_NEWTON : (opt_type == LM ? GPMP2B_OPT_LM : GPMP2B_OPT_DOGLEG);
    s.Qc = Qc.data(); s.opt_verbosity = opt_verbosity; s.final_iter_no_increase = final_iter_no_increase;
    s.rel_thresh = rel_thresh; s.max_iter = (int32_t)max_iter;
    if (goal_enabled) {
      s.goal_enabled = goal_is_pose ? 2 : 1; s.goal_link = goal_link;
      s.goal_keep_end_prior = goal_keep_end_prior; s.goal_sigma = goal_sigma;
      for (int k = 0; k < 9; k++) s.goal_R[k] = goal_R[k];
      for (int k = 0; k < 3; k++) s.goal_pos[k] = goal_pos[k];
      if (!goal_pos_batch.empty()) s.goal_pos_batch = goal_pos_batch.data();
      if (goal_is_pose && !goal_R_batch.empty()) s.goal_R_batch = goal_R_batch.data();
    }
    if (!self_collision_data.empty()) {
      if (self_collision_data.size() % 4) throw std::runtime_error("[TrajOptimizerSetting] ERROR: self-collision data must have 4 columns.");
      s.n_self_collision = (int32_t)(self_collision_data.size() / 4);
      s.self_collision_data = self_collision_data.data();
    }
    s.vehicle_dynamics_sigma = vehicle_dynamics_sigma;
    if (orient_enabled) {
      s.orient_enabled = 1; s.orient_link = orient_link; s.orient_sigma = orient_sigma;
      s.orient_state_first = orient_state_first; s.orient_state_last = orient_state_last < 0 ? (int32_t)total_step : orient_state_last;
      for (int k = 0; k < 9; k++) s.orient_R[k] = orient_R[k];
      if (!orient_R_batch.empty()) s.orient_R_batch = orient_R_batch.data();
    }
    if (fix_enabled) {
      s.fix_enabled = 1; s.fix_state_index = (int32_t)fix_state_index;
      s.fix_conf = fix_conf.data(); s.fix_vel = fix_vel.data();
    }
    return s;
  }
};

// ------------------------------------------------------------------------------------------------
/// per-problem results of a batched solve
struct BatchResult {
  std::vector<double> traj;        // [B][2*N*D] wire layout [x_0..x_T | v_0..v_T]
  std::vector<double> error, coll_cost;
  std::vector<int32_t> iters, status;
};

namespace detail {
inline Vector values_to_traj(const Values& v, size_t total_step, size_t D) {
  const size_t N = total_step + 1;
  Vector t(2 * N * D);
  for (size_t i = 0; i < N; i++) {
    const Vector& x = v.at(Symbol('x', i));
    const Vector& vel = v.at(Symbol('v', i));
    if (x.size() != D || vel.size() != D) throw std::runtime_error("init_values: dimension does not fit dof");
    for (size_t d = 0; d < D; d++) { t[i * D + d] = x[d]; t[(N + i) * D + d] = vel[d]; }
  }
  return t;
}
inline Values traj_to_values(const double* t, size_t total_step, size_t D) {
  const size_t N = total_step + 1;
  Values v;
  for (size_t i = 0; i < N; i++) {
    v.insert(Symbol('x', i), Vector(t + i * D, t + (i + 1) * D));
    v.insert(Symbol('v', i), Vector(t + (N + i) * D, t + (N + i + 1) * D));
  }
  return v;
}
template <class MODEL, class SDF>
BatchResult batch(const MODEL& model, const SDF& sdf, size_t B, const double* sc, const double* sv, const double* ec,
                  const double* ev, const double* init, const TrajOptimizerSetting& setting) {
  const size_t TL = 2 * (setting.total_step + 1) * setting.dof;
  BatchResult r;
  r.traj.resize(B * TL); r.error.resize(B); r.coll_cost.resize(B); r.iters.resize(B); r.status.resize(B);
  if ((!setting.goal_pos_batch.empty() && setting.goal_pos_batch.size() != 3 * B) || (!setting.goal_R_batch.empty() && setting.goal_R_batch.size() != 9 * B) ||
      (!setting.orient_R_batch.empty() && setting.orient_R_batch.size() != 9 * B) ||
      (setting.fix_enabled && (setting.fix_conf.size() != setting.dof * B || setting.fix_vel.size() != setting.dof * B)))
    throw std::runtime_error("per-problem workspace targets: the number of rows does not match the batch size");
  const gpmp2b_setting s = setting.pack();
  check(context(), gpmp2b_batch_optimize(context(), model.device(), sdf.device(), &s, (int64_t)B, sc, sv, ec, ev, init,
                                         r.traj.data(), r.error.data(), r.coll_cost.data(), r.iters.data(),
                                         r.status.data(), GPMP2B_MEM_HOST, nullptr));
  return r;
}
template <class MODEL, class SDF>
double collision_cost(const MODEL& model, const SDF& sdf, const Values& result, const TrajOptimizerSetting& setting) {
  // like the reference, i < result.size()/2 states (BatchTrajOptimizer-inl.h:96)
  TrajOptimizerSetting st = setting;
  st.total_step = result.size() / 2 - 1;
  const Vector t = values_to_traj(result, st.total_step, st.dof);
  const gpmp2b_setting s = st.pack();
  double c = 0.0;
  check(context(), gpmp2b_collision_cost(context(), model.device(), sdf.device(), &s, 1, t.data(), &c, GPMP2B_MEM_HOST, nullptr));
  return c;
}
/// RobotModel::sphereCentersMat on the device: gpmp2b_obstacle_errors over a two-state trajectory that holds the
/// configuration twice (no interpolated checks, a dummy 2 x 2 x 2 field; the errors are dropped, the centres returned)
inline Matrix sphere_centers_mat(const gpmp2b_robot* robot, size_t S, const Vector& q) {
  static const SignedDistanceField dummy(Point3(), 1.0, 2, 2, 2);
  const size_t D = q.size();
  TrajOptimizerSetting st(D);
  st.total_step = 1;
  st.obs_check_inter = 0;
  Vector traj(4 * D, 0.0);
  for (size_t k = 0; k < D; k++) traj[k] = traj[D + k] = q[k];
  std::vector<double> err(2 * S), ctr(2 * S * 3);
  const gpmp2b_setting s = st.pack();
  check(context(), gpmp2b_obstacle_errors(context(), robot, dummy.device(), &s, 1, traj.data(), err.data(), ctr.data(), GPMP2B_MEM_HOST, nullptr));
  Matrix m(3, S);
  for (size_t i = 0; i < S; i++)
    for (size_t k = 0; k < 3; k++) m(k, i) = ctr[i * 3 + k];
  return m;
}
}  // namespace detail

inline Matrix ArmModel::sphereCentersMat(const Vector& conf) const { return detail::sphere_centers_mat(device(), nr_body_spheres(), conf); }
inline Matrix Pose2MobileArmModel::sphereCentersMat(const Pose2Vector& conf) const {
  return detail::sphere_centers_mat(device(), nr_body_spheres(), conf.flat());
}
template <class FK>
Matrix MobileModelT<FK>::sphereCentersMat(const Pose2Vector& conf) const {
  return detail::sphere_centers_mat(device(), nr_body_spheres(), conf.flat());
}

// ---- batched overloads: B problems, flat row-major arrays [B][D] / [B][2*N*D] ----
inline BatchResult BatchTrajOptimize2DArm(const ArmModel& arm, const PlanarSDF& sdf, size_t B, const double* start_conf,
                                          const double* start_vel, const double* end_conf, const double* end_vel,
                                          const double* init_traj, const TrajOptimizerSetting& setting) {
  return detail::batch(arm, sdf, B, start_conf, start_vel, end_conf, end_vel, init_traj, setting);
}
inline BatchResult BatchTrajOptimize3DArm(const ArmModel& arm, const SignedDistanceField& sdf, size_t B,
                                          const double* start_conf, const double* start_vel, const double* end_conf,
                                          const double* end_vel, const double* init_traj, const TrajOptimizerSetting& setting) {
  return detail::batch(arm, sdf, B, start_conf, start_vel, end_conf, end_vel, init_traj, setting);
}

// ---- the reference's signatures (gpmp2/planner/BatchTrajOptimizer.h:43-66) ----
inline Values BatchTrajOptimize2DArm(const ArmModel& arm, const PlanarSDF& sdf, const Vector& start_conf,
                                     const Vector& start_vel, const Vector& end_conf, const Vector& end_vel,
                                     const Values& init_values, const TrajOptimizerSetting& setting) {
  const Vector t0 = detail::values_to_traj(init_values, setting.total_step, setting.dof);
  const BatchResult r = detail::batch(arm, sdf, 1, start_conf.data(), start_vel.data(), end_conf.data(), end_vel.data(), t0.data(), setting);
  return detail::traj_to_values(r.traj.data(), setting.total_step, setting.dof);
}
inline Values BatchTrajOptimize3DArm(const ArmModel& arm, const SignedDistanceField& sdf, const Vector& start_conf,
                                     const Vector& start_vel, const Vector& end_conf, const Vector& end_vel,
                                     const Values& init_values, const TrajOptimizerSetting& setting) {
  const Vector t0 = detail::values_to_traj(init_values, setting.total_step, setting.dof);
  const BatchResult r = detail::batch(arm, sdf, 1, start_conf.data(), start_vel.data(), end_conf.data(), end_vel.data(), t0.data(), setting);
  return detail::traj_to_values(r.traj.data(), setting.total_step, setting.dof);
}
/// Pose2Vector planners (BatchTrajOptimizer.h:81-104); Pose2Vector values travel flat as (x, y, theta, q...)
inline Values BatchTrajOptimizePose2MobileArm2D(const Pose2MobileArmModel& marm, const PlanarSDF& sdf,
                                                const Pose2Vector& start_conf, const Vector& start_vel,
                                                const Pose2Vector& end_conf, const Vector& end_vel,
                                                const Values& init_values, const TrajOptimizerSetting& setting) {
  const Vector t0 = detail::values_to_traj(init_values, setting.total_step, setting.dof);
  const Vector sc = start_conf.flat(), ec = end_conf.flat();
  const BatchResult r = detail::batch(marm, sdf, 1, sc.data(), start_vel.data(), ec.data(), end_vel.data(), t0.data(), setting);
  return detail::traj_to_values(r.traj.data(), setting.total_step, setting.dof);
}
inline Values BatchTrajOptimizePose2MobileArm(const Pose2MobileArmModel& marm, const SignedDistanceField& sdf,
                                              const Pose2Vector& start_conf, const Vector& start_vel,
                                              const Pose2Vector& end_conf, const Vector& end_vel,
                                              const Values& init_values, const TrajOptimizerSetting& setting) {
  const Vector t0 = detail::values_to_traj(init_values, setting.total_step, setting.dof);
  const Vector sc = start_conf.flat(), ec = end_conf.flat();
  const BatchResult r = detail::batch(marm, sdf, 1, sc.data(), start_vel.data(), ec.data(), end_vel.data(), t0.data(), setting);
  return detail::traj_to_values(r.traj.data(), setting.total_step, setting.dof);
}

/// the other Pose2Vector planners (BatchTrajOptimizer.h:106-125, BatchTrajOptimizer.cpp:92-128)
#define GPMP2B_MOBILE_PLANNER(NAME, MODEL)                                                                                  \
  inline Values BatchTrajOptimize##NAME(const MODEL& marm, const SignedDistanceField& sdf, const Pose2Vector& start_conf,    \
                                        const Vector& start_vel, const Pose2Vector& end_conf, const Vector& end_vel,         \
                                        const Values& init_values, const TrajOptimizerSetting& setting) {                    \
    const Vector t0 = detail::values_to_traj(init_values, setting.total_step, setting.dof);                                  \
    const Vector sc = start_conf.flat(), ec = end_conf.flat();                                                               \
    const BatchResult r = detail::batch(marm, sdf, 1, sc.data(), start_vel.data(), ec.data(), end_vel.data(), t0.data(), setting); \
    return detail::traj_to_values(r.traj.data(), setting.total_step, setting.dof);                                           \
  }                                                                                                                          \
  inline double CollisionCost##NAME(const MODEL& marm, const SignedDistanceField& sdf, const Values& result,                 \
                                    const TrajOptimizerSetting& setting) {                                                   \
    return detail::collision_cost(marm, sdf, result, setting);                                                               \
  }

// ---- CollisionCost* (BatchTrajOptimizer.h:135-185) ----
inline double CollisionCost2DArm(const ArmModel& arm, const PlanarSDF& sdf, const Values& result, const TrajOptimizerSetting& setting) {
  return detail::collision_cost(arm, sdf, result, setting);
}
inline double CollisionCost3DArm(const ArmModel& arm, const SignedDistanceField& sdf, const Values& result, const TrajOptimizerSetting& setting) {
  return detail::collision_cost(arm, sdf, result, setting);
}
inline double CollisionCostPose2MobileArm2D(const Pose2MobileArmModel& marm, const PlanarSDF& sdf, const Values& result, const TrajOptimizerSetting& setting) {
  return detail::collision_cost(marm, sdf, result, setting);
}
inline double CollisionCostPose2MobileArm(const Pose2MobileArmModel& marm, const SignedDistanceField& sdf, const Values& result, const TrajOptimizerSetting& setting) {
  return detail::collision_cost(marm, sdf, result, setting);
}
GPMP2B_MOBILE_PLANNER(Pose2Mobile2Arms, Pose2Mobile2ArmsModel)
GPMP2B_MOBILE_PLANNER(Pose2MobileVetLinArm, Pose2MobileVetLinArmModel)
GPMP2B_MOBILE_PLANNER(Pose2MobileVetLin2Arms, Pose2MobileVetLin2ArmsModel)
#undef GPMP2B_MOBILE_PLANNER

/// initArmTrajStraightLine, gpmp2/planner/TrajUtils.cpp:25-50 (avg_vel = (end - init) / total_step)
inline Values initArmTrajStraightLine(const Vector& init_conf, const Vector& end_conf, size_t total_step) {
  Values init_values;
  const size_t D = init_conf.size();
  for (size_t i = 0; i <= total_step; i++) {
    Vector conf(D);
    if (i == 0) conf = init_conf;
    else if (i == total_step) conf = end_conf;
    else
      for (size_t d = 0; d < D; d++)
        conf[d] = static_cast<double>(i) / static_cast<double>(total_step) * end_conf[d] +
                  (1.0 - static_cast<double>(i) / static_cast<double>(total_step)) * init_conf[d];
    init_values.insert(Symbol('x', i), conf);
  }
  Vector avg_vel(D);
  for (size_t d = 0; d < D; d++) avg_vel[d] = (end_conf[d] - init_conf[d]) / static_cast<double>(total_step);
  for (size_t i = 0; i <= total_step; i++) init_values.insert(Symbol('v', i), avg_vel);
  return init_values;
}

/// initPose2VectorTrajStraightLine, gpmp2/planner/TrajUtils.cpp:51-73 (device: gpmp2b_init_straight_line)
inline Values initPose2VectorTrajStraightLine(const Pose2& init_pose, const Vector& init_conf, const Pose2& end_pose,
                                              const Vector& end_conf, size_t total_step) {
  const Vector s = Pose2Vector(init_pose, init_conf).flat(), e = Pose2Vector(end_pose, end_conf).flat();
  if (s.size() != e.size()) throw std::runtime_error("initPose2VectorTrajStraightLine: dimension mismatch");
  Vector t(2 * (total_step + 1) * s.size());
  detail::check(detail::context(), gpmp2b_init_straight_line(detail::context(), GPMP2B_ROBOT_POSE2_MOBILE_ARM, (int)s.size(),
                                                             (int)total_step, 1, s.data(), e.data(), t.data(), GPMP2B_MEM_HOST, nullptr));
  return detail::traj_to_values(t.data(), total_step, s.size());
}

namespace detail {
inline Values interpolate(int kind, const Values& opt_values, const double* Qc, double delta_t, size_t inter_step,
                          size_t start_index, size_t end_index) {
  size_t total_step = 0;
  while (opt_values.exists(Symbol('x', total_step + 1))) total_step++;
  const size_t D = opt_values.at(Symbol('x', 0)).size();
  const Vector t = values_to_traj(opt_values, total_step, D);
  const size_t nout = (end_index - start_index) * (inter_step + 1) + 1;
  Vector out(2 * nout * D);
  check(context(), gpmp2b_interpolate_traj(context(), kind, (int)D, (int)total_step, delta_t, Qc, (int)inter_step,
                                           (int)start_index, (int)end_index, 1, t.data(), out.data(), GPMP2B_MEM_HOST, nullptr));
  return traj_to_values(out.data(), nout - 1, D);
}
inline size_t count_steps(const Values& v) {
  size_t n = 0;
  while (v.exists(Symbol('x', n + 1))) n++;
  return n;
}
}  // namespace detail

/// interpolateArmTraj, gpmp2/planner/TrajUtils.cpp:96-196.  Qc: row-major dof x dof covariance (the reference's
/// Qc_model), or nullptr for identity.
inline Values interpolateArmTraj(const Values& opt_values, const double* Qc, double delta_t, size_t inter_step) {
  return detail::interpolate(GPMP2B_ROBOT_ARM, opt_values, Qc, delta_t, inter_step, 0, detail::count_steps(opt_values));
}
inline Values interpolateArmTraj(const Values& opt_values, const double* Qc, double delta_t, size_t inter_step,
                                 size_t start_index, size_t end_index) {
  return detail::interpolate(GPMP2B_ROBOT_ARM, opt_values, Qc, delta_t, inter_step, start_index, end_index);
}
/// interpolatePose2MobileArmTraj, gpmp2/planner/TrajUtils.cpp:199-237
inline Values interpolatePose2MobileArmTraj(const Values& opt_values, const double* Qc, double delta_t, size_t inter_step,
                                            size_t start_index, size_t end_index) {
  return detail::interpolate(GPMP2B_ROBOT_POSE2_MOBILE_ARM, opt_values, Qc, delta_t, inter_step, start_index, end_index);
}

/// initPose2TrajStraightLine (gpmp2/planner/TrajUtils.cpp:76-93) and interpolatePose2Traj (:239-275): a bare Pose2 trajectory is
/// the Pose2Vector one with an empty arm (dof 3; GaussianProcessInterpolatorPose2 is the same Lie interpolator on SE(2)); the
/// values under x(i) are (x, y, theta)
inline Values initPose2TrajStraightLine(const Pose2& init_pose, const Pose2& end_pose, size_t total_step) {
  return initPose2VectorTrajStraightLine(init_pose, Vector(), end_pose, Vector(), total_step);
}
inline Values interpolatePose2Traj(const Values& opt_values, const double* Qc, double delta_t, size_t inter_step,
                                   size_t start_index, size_t end_index) {
  return detail::interpolate(GPMP2B_ROBOT_POSE2_MOBILE_ARM, opt_values, Qc, delta_t, inter_step, start_index, end_index);
}

/// best of `restarts` consecutive problems per query of a BatchResult (gpmp2b_select_best): indices into the batch
inline std::vector<int64_t> selectBest(const std::vector<double>& error, const std::vector<double>& coll_cost,
                                       size_t restarts, double coll_tol = 0.0, std::vector<int32_t>* feasible = nullptr) {
  if (restarts == 0 || error.size() % restarts || (!coll_cost.empty() && coll_cost.size() != error.size()))
    throw std::runtime_error("selectBest: sizes do not fit");
  const size_t G = error.size() / restarts;
  std::vector<int64_t> best(G);
  std::vector<int32_t> feas(G);
  detail::check(detail::context(), gpmp2b_select_best(detail::context(), (int64_t)G, (int64_t)restarts, error.data(),
                                                      coll_cost.empty() ? nullptr : coll_cost.data(), coll_tol, best.data(),
                                                      feas.data(), GPMP2B_MEM_HOST, nullptr));
  if (feasible) *feasible = feas;
  return best;
}

}  // namespace gpmp2
