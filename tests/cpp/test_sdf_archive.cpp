// SignedDistanceField::saveSDF / loadSDF of the C++ facade (include/gpmp2b/gpmp2.hpp) -- host-side only, no device call.
//   test_sdf_archive rewrite <in> <out>   load <in>, print the header, save as <out>
//   test_sdf_archive values               the Pose2Vector-in-Values helpers of the wrapper interface
//   test_sdf_archive make <out>           build a small field with the reference's constructor + initFieldData and save it
#include <cstdio>
#include <string>

#include "gpmp2b/gpmp2.hpp"

int main(int argc, char** argv) {
  try {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "rewrite" && argc == 4) {
      gpmp2::SignedDistanceField sdf(gpmp2::Point3(), 1.0, 1, 1, 1);
      sdf.loadSDF(argv[2]);
      std::printf("%zu %zu %zu %.17g %.17g %.17g %.17g\n", sdf.y_count(), sdf.x_count(), sdf.z_count(), sdf.cell_size(),
                  sdf.origin().x(), sdf.origin().y(), sdf.origin().z());
      sdf.saveSDF(argv[3]);
      return 0;
    }
    if (mode == "make" && argc == 3) {
      gpmp2::SignedDistanceField sdf(gpmp2::Point3(-0.2, 0.4, 1.0 / 3.0), 0.01, 3, 2, 4);   // rows, cols, z
      for (size_t z = 0; z < 4; z++) {
        gpmp2::Matrix m(3, 2);
        for (size_t r = 0; r < 3; r++)
          for (size_t c = 0; c < 2; c++) m(r, c) = (100.0 * z + 10.0 * r + c) / 7.0;
        sdf.initFieldData(z, m);
      }
      sdf.saveSDF(argv[2]);
      return 0;
    }
    if (mode == "values") {   // insertPose2VectorInValues / atPose2VectorValues (gpmp2/utils/matlabUtils.cpp:14-22)
      gpmp2::Values v;
      gpmp2::insertPose2VectorInValues(gpmp2::Symbol('x', 3), gpmp2::Pose2Vector(gpmp2::Pose2(1.0, -2.0, 0.5), gpmp2::Vector{0.25, 0.75}), v);
      const gpmp2::Pose2Vector p = gpmp2::atPose2VectorValues(gpmp2::Symbol('x', 3), v);
      std::printf("%g %g %g %zu %g %g\n", p.pose().x(), p.pose().y(), p.pose().theta(), p.configuration().size(),
                  p.configuration()[0], p.configuration()[1]);
      try { gpmp2::atPose2VectorValues(gpmp2::Symbol('x', 4), v); } catch (const std::runtime_error& e) { std::printf("%s\n", e.what()); }
      try { gpmp2::insertPose2VectorInValues(gpmp2::Symbol('x', 3), p, v); } catch (const std::runtime_error& e) { std::printf("%s\n", e.what()); }
      return 0;
    }
    if (mode == "extras") {   // compiles everywhere, runs on a GPU box only: sphereCentersMat, Pose2 trajectory utilities
      gpmp2::Arm arm(2, gpmp2::Vector{1.0, 1.0}, gpmp2::Vector{0.0, 0.0}, gpmp2::Vector{0.0, 0.0});
      gpmp2::BodySphereVector sph;
      sph.push_back(gpmp2::BodySphere(0, 0.1, gpmp2::Point3(-0.5, 0, 0)));
      sph.push_back(gpmp2::BodySphere(1, 0.1, gpmp2::Point3(0, 0, 0)));
      const gpmp2::Matrix m = gpmp2::ArmModel(arm, sph).sphereCentersMat(gpmp2::Vector{0.0, 1.5707963267948966});
      std::printf("centers %.12g %.12g %.12g | %.12g %.12g %.12g\n", m(0, 0), m(1, 0), m(2, 0), m(0, 1), m(1, 1), m(2, 1));
      const gpmp2::Values v = gpmp2::initPose2TrajStraightLine(gpmp2::Pose2(0, 0, 0), gpmp2::Pose2(2, 4, 1), 4);
      const gpmp2::Values d = gpmp2::interpolatePose2Traj(v, nullptr, 0.5, 1, 0, 4);
      std::printf("pose2 %zu %zu %.12g %.12g %.12g\n", v.size(), d.size(), v.at(gpmp2::Symbol('x', 2))[0], v.at(gpmp2::Symbol('x', 2))[1],
                  v.at(gpmp2::Symbol('v', 0))[2]);
      return 0;
    }
    std::fprintf(stderr, "usage: test_sdf_archive rewrite <in> <out> | make <out>\n");
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
