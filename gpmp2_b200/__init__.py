"""gpmp2_b200 -- B200-native batched GPMP2 trajectory-optimization hot path.

Host-side mirror of the reference's planner interface (api.py) over the C ABI in
include/gpmp2b.h, implemented by hand-written sm_100a CUDA in csrc/.  No CPU fallback.
"""
from . import _abi  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (Arm, ArmModel, BodySphere, Context, PlanarSDF, Pose2, Pose2MobileArm,  # noqa: F401
                  Pose2MobileArmModel, Pose2Vector, Pose3, SignedDistanceField, TrajOptimizerSetting,
                  Values, batch_collision_cost, batch_linearize, batch_obstacle_errors, batch_optimize,
                  default_context, initArmTrajStraightLine, Pose2Mobile2Arms, Pose2Mobile2ArmsModel,
                  Pose2MobileVetLinArm, Pose2MobileVetLinArmModel, Pose2MobileVetLin2Arms, Pose2MobileVetLin2ArmsModel, readSDFvolfile, straight_line_traj, symbol, writeSDFvolfile,
                  insertPose2VectorInValues, atPose2VectorValues,
                  initPose2TrajStraightLine, interpolatePose2Traj)
