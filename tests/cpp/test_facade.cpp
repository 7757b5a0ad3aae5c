// C++ facade test, written like the reference's own unit tests (fixtures from
// gpmp2/obstacle/tests/testObstaclePlanarSDFFactorArm.cpp:36-108 and a WAM planning call in the style of
// matlab/WAMPlannerExample.m).  Prints results as "key value..." lines; tests/test_cpp_facade.py compares them
// with the oracle.  Without a CUDA device it must fail loudly (exit code 3), never fall back.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "gpmp2b/gpmp2.hpp"

using namespace gpmp2;

int main(int argc, char** argv) {
  try {
    // 2 link simple example
    Pose3 arm_base = Pose3::Translation(Point3(0.5, 1.5, 0));
    Arm abs_arm(2, Vector{1, 2}, Vector{0, 0}, Vector{0, 0}, arm_base);
    BodySphereVector body_spheres;
    const double r = 0.5;
    body_spheres.push_back(BodySphere(0, r, Point3(-1, 0, 0)));
    body_spheres.push_back(BodySphere(0, r, Point3(0, 0, 0)));
    body_spheres.push_back(BodySphere(1, r, Point3(-1, 0, 0)));
    body_spheres.push_back(BodySphere(1, r, Point3(0, 0, 0)));
    ArmModel arm(abs_arm, body_spheres);

    const double f[7][7] = {{2.8284, 2.2361, 2.0000, 2.0000, 2.0000, 2.2361, 2.8284}, {2.2361, 1.4142, 1.0000, 1.0000, 1.0000, 1.4142, 2.2361},
                            {2.0000, 1.0000, -1.0000, -1.0000, -1.0000, 1.0000, 2.0000}, {2.0000, 1.0000, -1.0000, -2.0000, -1.0000, 1.0000, 2.0000},
                            {2.0000, 1.0000, -1.0000, -1.0000, -1.0000, 1.0000, 2.0000}, {2.2361, 1.4142, 1.0000, 1.0000, 1.0000, 1.4142, 2.2361},
                            {2.8284, 2.2361, 2.0000, 2.0000, 2.0000, 2.2361, 2.8284}};
    Matrix field(7, 7);
    for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) field(i, j) = f[i][j];
    PlanarSDF sdf(Point2(0, 0), 1.0, field);

    TrajOptimizerSetting setting(2);
    setting.set_total_step(4);
    setting.set_total_time(2.0);
    setting.set_epsilon(1.0);
    setting.set_cost_sigma(0.5);
    setting.set_obs_check_inter(2);
    setting.setLM();
    setting.set_max_iter(8);
    setting.set_rel_thresh(1e-6);

    Vector start_conf{0, 0}, end_conf{M_PI / 2, 0}, zero{0, 0};
    Values init_values = initArmTrajStraightLine(start_conf, end_conf, setting.total_step);
    Values result = BatchTrajOptimize2DArm(arm, sdf, start_conf, zero, end_conf, zero, init_values, setting);
    if (result.size() != 2 * (setting.total_step + 1)) throw std::runtime_error("wrong number of values");
    for (size_t i = 0; i <= setting.total_step; i++) {
      const Vector& x = result.at(Symbol('x', i));
      const Vector& v = result.at(Symbol('v', i));
      std::printf("x%zu %.17g %.17g\nv%zu %.17g %.17g\n", i, x[0], x[1], i, v[0], v[1]);
    }
    std::printf("coll_cost %.17g\n", CollisionCost2DArm(arm, sdf, result, setting));

    // the optional factors of hand-built graphs through the facade's setters: pose goal on x_T instead
    // of the end prior, self-collision pairs on every state, orientation prior on the interior states
    {
      TrajOptimizerSetting ex = setting;
      const double Rg[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
      ex.set_workspace_pose_goal(Rg, 0.6, 3.4, 0.0, 0.2);
      ex.set_self_collision(Vector{0, 3, 0.2, 0.5, 1, 2, 0.1, 0.3});
      const double Ro[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      ex.set_workspace_orientation(Ro, 2.0, 0, 1, 3);
      Values rex = BatchTrajOptimize2DArm(arm, sdf, start_conf, zero, end_conf, zero, init_values, ex);
      for (size_t i = 0; i <= ex.total_step; i++) {
        const Vector& x = rex.at(Symbol('x', i));
        std::printf("ex%zu %.17g %.17g\n", i, x[0], x[1]);
      }
    }

    // error behaviour mirrors the reference's exceptions
    bool threw = false;
    try {
      TrajOptimizerSetting bad(2);
      bad.setLM();
      bad.set_flag_vel_limit(true);
      bad.set_vel_limits(Vector{0.0, 1.0});   // VelocityLimitFactorVector.h:54-56
      BatchTrajOptimize2DArm(arm, sdf, start_conf, zero, end_conf, zero, initArmTrajStraightLine(start_conf, end_conf, 10), bad);
    } catch (const std::runtime_error& e) {
      threw = true;
      std::printf("threw %s\n", e.what());
    }
    if (!threw) throw std::runtime_error("expected std::runtime_error for velocity limit <= 0");
    try { result.at(Symbol('x', 99)); threw = false; } catch (const std::runtime_error&) { threw = true; }
    if (!threw) throw std::runtime_error("expected ValuesKeyDoesNotExist");
    // trajectory utilities (gpmp2/planner/tests/testTrajUtils.cpp:28-53, 56-62)
    {
      Values iv;
      iv.insert(Symbol('x', 0), Vector{0, 0}); iv.insert(Symbol('x', 1), Vector{1, 0});
      iv.insert(Symbol('v', 0), Vector{10, 0}); iv.insert(Symbol('v', 1), Vector{10, 0});
      const double Qc[4] = {0.01, 0, 0, 0.01};
      const Values inter = interpolateArmTraj(iv, Qc, 0.1, 4);
      for (size_t k = 0; k <= 5; k++) {
        const Vector& x = inter.at(Symbol('x', k));
        const Vector& v = inter.at(Symbol('v', k));
        if (std::fabs(x[0] - 0.2 * k) > 1e-6 || std::fabs(x[1]) > 1e-6 || std::fabs(v[0] - 10) > 1e-6 || std::fabs(v[1]) > 1e-6)
          throw std::runtime_error("interpolateArmTraj golden mismatch");
      }
      const Values pv = initPose2VectorTrajStraightLine(Pose2(1, 3, M_PI - 0.5), Vector{2, 4}, Pose2(3, 7, -M_PI + 0.5), Vector{4, 8}, 5);
      const Vector& x0 = pv.at(Symbol('x', 0));
      const double want[5] = {1, 3, M_PI - 0.5, 2, 4};
      for (int d = 0; d < 5; d++) if (std::fabs(x0[d] - want[d]) > 1e-6) throw std::runtime_error("initPose2VectorTrajStraightLine golden mismatch");
      std::vector<int32_t> feas;
      const std::vector<int64_t> best = selectBest({3.0, 1.0, 2.0, 5.0, 4.0, 6.0}, {0.0, 1.0, 0.0, 1.0, 1.0, 1.0}, 3, 0.0, &feas);
      if (best.size() != 2 || best[0] != 2 || feas[0] != 1 || best[1] != 4 || feas[1] != 0) throw std::runtime_error("selectBest mismatch");
      // signed distance field from an occupancy grid: one obstacle cell in a 5 x 6 map, distances in cells * 0.5
      Matrix occ(5, 6);
      for (size_t r = 0; r < 5; r++) for (size_t c = 0; c < 6; c++) occ(r, c) = 0.0;
      occ(2, 3) = 1.0;
      const Matrix f2 = signedDistanceField2D(occ, 0.5, false);
      if (std::fabs(f2(2, 3) + 0.5) > 1e-12 || std::fabs(f2(0, 0) - 0.5 * std::sqrt(13.0)) > 1e-12 || std::fabs(f2(2, 5) - 1.0) > 1e-12)
        throw std::runtime_error("signedDistanceField2D mismatch");
      std::printf("trajutils ok\n");
    }
    // the other Pose2Vector robots (BatchTrajOptimizer.h:106-125): a VetLin2Arms planning call in an empty field keeps the
    // straight line; Pose2Mobile2Arms / Pose2MobileVetLinArm models upload
    {
      Arm a2(2, Vector{0.3, 0.3}, Vector{0, 0}, Vector{0, 0}), a1(1, Vector{0.35}, Vector{0}, Vector{0});
      BodySphereVector sph;
      for (size_t l = 0; l < 5; l++) sph.push_back(BodySphere(l, 0.05, Point3(0, 0, 0)));
      Pose2MobileVetLin2ArmsModel m(Pose2MobileVetLin2Arms(a2, a1, Pose3(), Pose3::Translation(Point3(0.1, 0.1, 0)), Pose3::Translation(Point3(0.1, -0.1, 0)), true), sph);
      BodySphereVector sph4(sph.begin(), sph.begin() + 4);
      Pose2Mobile2ArmsModel m2(Pose2Mobile2Arms(a2, a1), sph4);
      Pose2MobileVetLinArmModel m3(Pose2MobileVetLinArm(a2), sph4);
      if (m.dof() != 7 || m2.dof() != 6 || m3.dof() != 6) throw std::runtime_error("mobile robot dof");
      std::vector<Matrix> layers(4, Matrix(6, 6));
      for (auto& L : layers) for (size_t r = 0; r < 6; r++) for (size_t c = 0; c < 6; c++) L(r, c) = 10.0;
      SignedDistanceField far(Point3(-3, -3, -1), 1.5, layers);
      TrajOptimizerSetting st(7);
      st.set_total_step(4); st.set_total_time(2.0); st.set_obs_check_inter(1); st.setLM(); st.set_max_iter(5);
      const Pose2Vector ps(Pose2(0, 0, 0), Vector{0.1, 0, 0, 0}), pe(Pose2(1, 0.5, 0.4), Vector{0.3, 0.2, -0.2, 0.1});
      const Values iv = initPose2VectorTrajStraightLine(ps.pose(), ps.configuration(), pe.pose(), pe.configuration(), 4);
      const Vector z7(7, 0.0);
      const Values res = BatchTrajOptimizePose2MobileVetLin2Arms(m, far, ps, z7, pe, z7, iv, st);
      const Vector& xe = res.at(Symbol('x', 4));
      if (std::fabs(xe[0] - 1.0) > 1e-3 || std::fabs(xe[3] - 0.3) > 1e-3) throw std::runtime_error("VetLin2Arms planning call");
      if (CollisionCostPose2MobileVetLin2Arms(m, far, res, st) != 0.0) throw std::runtime_error("VetLin2Arms collision cost");
      std::printf("othermobile ok\n");
    }
    std::printf("ok\n");
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "test_facade: %s\n", e.what());
    return 3;
  }
}
