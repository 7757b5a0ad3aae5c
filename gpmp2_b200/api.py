"""Host-side mirror of the reference's planner interface for the batched trajectory-optimization path.

Same names, argument meaning and error behaviour as the reference (ori-drs/gpmp2) so tests read like
the reference's own:

  Arm(dof, a, alpha, d[, base_pose[, theta_bias]])      gpmp2/kinematics/Arm.h:47-55
  BodySphere(id, r, center), ArmModel(arm, spheres)     gpmp2/kinematics/RobotModel.h:20-27,56
  Pose2MobileArm(arm, base_T_arm), Pose2MobileArmModel  gpmp2/kinematics/Pose2MobileArm.h:41
  PlanarSDF(origin, cell_size, data)                    gpmp2/obstacle/PlanarSDF.h:45-47
  SignedDistanceField(origin, cell_size, data) / (origin, cell, rows, cols, z) + initFieldData
                                                        gpmp2/obstacle/SignedDistanceField.h:58-79
  TrajOptimizerSetting(dof) + setters                   gpmp2/planner/TrajOptimizerSetting.h:17-100
  BatchTrajOptimize2DArm / 3DArm / Pose2MobileArm2D / Pose2MobileArm, CollisionCost*
                                                        gpmp2/planner/BatchTrajOptimizer.h:43-185
  initArmTrajStraightLine                               gpmp2/planner/TrajUtils.cpp:25-50

This module only packs arguments for the C ABI (include/gpmp2b.h); all arithmetic of the hot path
runs in the CUDA library.  There is no CPU fallback.
"""
import ctypes as C
import itertools
import weakref

import numpy as np

from . import _abi


# ------------------------------------------------------------------------------------------------
# light geometry value types (the reference uses gtsam::Pose3 / Pose2 / Point3)
# ------------------------------------------------------------------------------------------------
class Pose3:
    """Rigid transform held as a 4x4 matrix (gtsam::Pose3::matrix())."""

    def __init__(self, R=None, t=None, matrix=None):
        if matrix is not None:
            self.T = np.array(matrix, dtype=np.float64).reshape(4, 4)
        else:
            self.T = np.eye(4)
            if R is not None:
                self.T[:3, :3] = np.asarray(R, dtype=np.float64)
            if t is not None:
                self.T[:3, 3] = np.asarray(t, dtype=np.float64)

    def matrix(self):
        return self.T


class Pose2:
    def __init__(self, x=0.0, y=0.0, theta=0.0):
        self._x, self._y, self._theta = float(x), float(y), float(theta)

    def x(self):
        return self._x

    def y(self):
        return self._y

    def theta(self):
        return self._theta


class Pose2Vector:
    """SE(2) x R^n state of a mobile manipulator (gpmp2/geometry/Pose2Vector.h:26-53)."""

    def __init__(self, pose, conf):
        self._pose = pose
        self._conf = np.asarray(conf, dtype=np.float64).ravel()

    def pose(self):
        return self._pose

    def configuration(self):
        return self._conf

    def flat(self):
        """(x, y, theta, q...) -- the wire layout of include/gpmp2b.h."""
        return np.concatenate([[self._pose.x(), self._pose.y(), self._pose.theta()], self._conf])

    @staticmethod
    def from_flat(v):
        return Pose2Vector(Pose2(v[0], v[1], v[2]), v[3:])


def symbol(c, i):
    """gtsam::Symbol(c, i) -> hashable key."""
    return (c, int(i))


class Values(dict):
    """Minimal stand-in for gtsam::Values: key -> vector or Pose2Vector."""

    def insert(self, key, value):
        if key in self:
            raise KeyError("ValuesKeyAlreadyExists: %r" % (key,))
        self[key] = value

    def atVector(self, key):
        if key not in self:
            raise KeyError("ValuesKeyDoesNotExist: %r" % (key,))
        return self[key]

    at = atVector

    def size(self):
        return len(self)


def insertPose2VectorInValues(key, p, values):
    """gpmp2::insertPose2VectorInValues (gpmp2/utils/matlabUtils.cpp:14-17): what the MATLAB / Python toolboxes call to put a
    Pose2Vector into the init_values of a mobile-manipulator planner."""
    if not isinstance(p, Pose2Vector):
        raise TypeError("insertPose2VectorInValues: p must be a Pose2Vector")
    values.insert(key, p)


def atPose2VectorValues(key, values):
    """gpmp2::atPose2VectorValues (gpmp2/utils/matlabUtils.cpp:20-22): values.at<Pose2Vector>(key).  The planners return
    Pose2Vector values; a flat (x, y, theta, q...) vector stored under the key is accepted too."""
    v = values.at(key)
    return v if isinstance(v, Pose2Vector) else Pose2Vector.from_flat(np.asarray(v, dtype=np.float64).ravel())


# ------------------------------------------------------------------------------------------------
# robot models
# ------------------------------------------------------------------------------------------------
class Arm:
    def __init__(self, dof, a, alpha, d, base_pose=None, theta_bias=None):
        self._dof = int(dof)
        self._a = np.asarray(a, dtype=np.float64).ravel().copy()
        self._alpha = np.asarray(alpha, dtype=np.float64).ravel().copy()
        self._d = np.asarray(d, dtype=np.float64).ravel().copy()
        for v in (self._a, self._alpha, self._d):
            if v.size != self._dof:
                raise RuntimeError("[Arm] ERROR: DH parameter dim does not fit dof.")
        self._base = base_pose if base_pose is not None else Pose3()
        self._bias = (np.zeros(self._dof) if theta_bias is None
                      else np.asarray(theta_bias, dtype=np.float64).ravel().copy())

    def dof(self):
        return self._dof

    def a(self):
        return self._a

    def d(self):
        return self._d

    def alpha(self):
        return self._alpha

    def base_pose(self):
        return self._base


class BodySphere:
    def __init__(self, link_id, radius, center):
        self.link_id = int(link_id)
        self.radius = float(radius)
        self.center = np.asarray(center, dtype=np.float64).ravel().copy()


class _RobotModelBase:
    def __del__(self):   # free the device copies with the object (the ctx would otherwise keep them until it closes)
        try:
            _drop_handles(self, "robot")
        except Exception:
            pass

    """Packs a gpmp2b_robot_desc; keeps the numpy buffers alive."""

    kind = _abi.ROBOT_ARM

    def _pack(self, arm, base_matrix, spheres):
        self._spheres = list(spheres)
        n = len(self._spheres)
        if n > _abi.MAX_SPHERES:
            raise RuntimeError("too many body spheres (max %d)" % _abi.MAX_SPHERES)
        self._link = np.array([s.link_id for s in self._spheres], dtype=np.int32)
        self._radius = np.array([s.radius for s in self._spheres], dtype=np.float64)
        self._center = np.ascontiguousarray(
            np.array([s.center for s in self._spheres], dtype=np.float64).reshape(n, 3))
        self._arm = arm
        d = _abi.RobotDesc()
        d.kind = self.kind
        d.arm_dof = arm.dof()
        d.n_spheres = n
        d.a, d.alpha, d.d = _abi.dptr(arm._a), _abi.dptr(arm._alpha), _abi.dptr(arm._d)
        d.theta_bias = _abi.dptr(arm._bias)
        d.base_pose = (C.c_double * 16)(*np.asarray(base_matrix, dtype=np.float64).reshape(16))
        d.sphere_link, d.sphere_radius, d.sphere_center = _abi.iptr(self._link), _abi.dptr(self._radius), _abi.dptr(self._center)
        self.desc = d
        self._handles = {}

    def nr_body_spheres(self):
        return len(self._spheres)

    def sphere_radius(self, i):
        return self._spheres[i].radius

    def sphere_link_id(self, i):
        return self._spheres[i].link_id

    def sphereCentersMat(self, conf, ctx=None):
        """RobotModel::sphereCentersMat (gpmp2/kinematics/RobotModel-inl.h:71-82): the 3 x S matrix of sphere centres in the
        world frame at one configuration (a vector, or a Pose2Vector for the mobile manipulators).  On the device, through
        gpmp2b_obstacle_errors: a two-state trajectory holding the configuration twice, no interpolated checks, a dummy
        field (the errors are ignored, the centres returned)."""
        q = _flat_state(conf)
        D = q.size
        st = TrajOptimizerSetting(D)
        st.set_total_step(1)
        st.set_obs_check_inter(0)
        out = batch_obstacle_errors(self, _dummy_field(), np.concatenate([q, q, np.zeros(2 * D)]), st, ctx=ctx)
        return np.ascontiguousarray(out["centers"][0, 0].T)


class ArmModel(_RobotModelBase):
    kind = _abi.ROBOT_ARM

    def __init__(self, arm, spheres):
        self._pack(arm, arm.base_pose().matrix(), spheres)
        for s in self._spheres:
            if not 0 <= s.link_id < arm.dof():
                raise RuntimeError("[ArmModel] sphere link id out of range")

    def dof(self):
        return self._arm.dof()

    def fk_model(self):
        return self._arm


class Pose2MobileArm:
    def __init__(self, arm, base_T_arm=None):
        self._arm = arm
        self._base_T_arm = base_T_arm if base_T_arm is not None else Pose3()

    def dof(self):
        return self._arm.dof() + 3

    def nr_links(self):
        return self._arm.dof() + 1

    def arm(self):
        return self._arm

    def base_T_arm(self):
        return self._base_T_arm


class Pose2MobileArmModel(_RobotModelBase):
    kind = _abi.ROBOT_POSE2_MOBILE_ARM

    def __init__(self, marm, spheres):
        self._marm = marm
        self._pack(marm.arm(), marm.base_T_arm().matrix(), spheres)
        for s in self._spheres:
            if not 0 <= s.link_id < marm.nr_links():
                raise RuntimeError("[Pose2MobileArmModel] sphere link id out of range")

    def dof(self):
        return self._marm.dof()

    def fk_model(self):
        return self._marm


class _TwoArms(Arm):
    """The DH tables of arm 1 followed by those of arm 2 (the layout gpmp2b_robot_desc takes for the two-arm robots)."""

    def __init__(self, arm1, arm2):
        cat = lambda f: np.ascontiguousarray(np.concatenate([f(arm1), f(arm2)]))
        self._dof = arm1.dof() + arm2.dof()
        self._a, self._alpha, self._d = cat(lambda a: a._a), cat(lambda a: a._alpha), cat(lambda a: a._d)
        self._bias = cat(lambda a: a._bias if a._bias is not None else np.zeros(a.dof()))
        self._base = arm1._base

    def dof(self):
        return self._dof


class Pose2Mobile2Arms:
    """gpmp2::Pose2Mobile2Arms (gpmp2/kinematics/Pose2Mobile2Arms.h:23-66): vehicle + two arms; state = Pose2Vector
    (x, y, theta | arm 1 joints | arm 2 joints); links 0 = vehicle, then arm 1's joint frames, then arm 2's."""

    def __init__(self, arm1, arm2, base_T_arm1=None, base_T_arm2=None):
        self._arm1, self._arm2 = arm1, arm2
        self._b1 = base_T_arm1 if base_T_arm1 is not None else Pose3()
        self._b2 = base_T_arm2 if base_T_arm2 is not None else Pose3()

    def dof(self): return self._arm1.dof() + self._arm2.dof() + 3
    def nr_links(self): return self._arm1.dof() + self._arm2.dof() + 1
    def arm1(self): return self._arm1
    def arm2(self): return self._arm2
    def base_T_arm1(self): return self._b1
    def base_T_arm2(self): return self._b2


class Pose2MobileVetLinArm:
    """gpmp2::Pose2MobileVetLinArm (gpmp2/kinematics/Pose2MobileVetLinArm.h:24-68): vehicle + vertical linear actuator +
    arm; state = Pose2Vector (x, y, theta | z | arm joints); links 0 = vehicle, 1 = torso, 2.. = arm joint frames."""

    def __init__(self, arm, base_T_torso=None, torso_T_arm=None, reverse_linact=False):
        self._arm = arm
        self._bt = base_T_torso if base_T_torso is not None else Pose3()
        self._ta = torso_T_arm if torso_T_arm is not None else Pose3()
        self._rev = bool(reverse_linact)

    def dof(self): return self._arm.dof() + 4
    def nr_links(self): return self._arm.dof() + 2
    def arm(self): return self._arm
    def base_T_torso(self): return self._bt
    def torso_T_arm(self): return self._ta
    def reverse_linact(self): return self._rev


class Pose2MobileVetLin2Arms:
    """gpmp2::Pose2MobileVetLin2Arms (gpmp2/kinematics/Pose2MobileVetLin2Arms.h:24-75): vehicle + linear actuator + two arms;
    state = Pose2Vector (x, y, theta | z | arm 1 joints | arm 2 joints); links 0 = vehicle, 1 = torso, arm 1, arm 2."""

    def __init__(self, arm1, arm2, base_T_torso=None, torso_T_arm1=None, torso_T_arm2=None, reverse_linact=False):
        self._arm1, self._arm2 = arm1, arm2
        self._bt = base_T_torso if base_T_torso is not None else Pose3()
        self._t1 = torso_T_arm1 if torso_T_arm1 is not None else Pose3()
        self._t2 = torso_T_arm2 if torso_T_arm2 is not None else Pose3()
        self._rev = bool(reverse_linact)

    def dof(self): return self._arm1.dof() + self._arm2.dof() + 4
    def nr_links(self): return self._arm1.dof() + self._arm2.dof() + 2
    def arm1(self): return self._arm1
    def arm2(self): return self._arm2
    def base_T_torso(self): return self._bt
    def torso_T_arm1(self): return self._t1
    def torso_T_arm2(self): return self._t2
    def reverse_linact(self): return self._rev


class _MobileModelBase(_RobotModelBase):
    def _finish(self, marm, arm2_dof=0, reverse=False, base2=None, base3=None):
        self._marm = marm
        self.desc.arm2_dof = int(arm2_dof)
        self.desc.reverse_linact = int(bool(reverse))
        if base2 is not None:
            self.desc.base_pose2 = (C.c_double * 16)(*np.asarray(base2.matrix(), dtype=np.float64).reshape(16))
        if base3 is not None:
            self.desc.base_pose3 = (C.c_double * 16)(*np.asarray(base3.matrix(), dtype=np.float64).reshape(16))
        for s in self._spheres:
            if not 0 <= s.link_id < marm.nr_links():
                raise RuntimeError("[%s] sphere link id out of range" % type(self).__name__)

    def dof(self):
        return self._marm.dof()

    def fk_model(self):
        return self._marm


class Pose2Mobile2ArmsModel(_MobileModelBase):
    kind = _abi.ROBOT_POSE2_MOBILE_2ARMS

    def __init__(self, marm, spheres):
        arms = _TwoArms(marm.arm1(), marm.arm2())
        self._pack(arms, marm.base_T_arm1().matrix(), spheres)
        self.desc.arm_dof = marm.arm1().dof()
        self._finish(marm, marm.arm2().dof(), False, marm.base_T_arm2())


class Pose2MobileVetLinArmModel(_MobileModelBase):
    kind = _abi.ROBOT_POSE2_MOBILE_VETLIN_ARM

    def __init__(self, marm, spheres):
        self._pack(marm.arm(), marm.base_T_torso().matrix(), spheres)
        self._finish(marm, 0, marm.reverse_linact(), marm.torso_T_arm())


class Pose2MobileVetLin2ArmsModel(_MobileModelBase):
    kind = _abi.ROBOT_POSE2_MOBILE_VETLIN_2ARMS

    def __init__(self, marm, spheres):
        arms = _TwoArms(marm.arm1(), marm.arm2())
        self._pack(arms, marm.base_T_torso().matrix(), spheres)
        self.desc.arm_dof = marm.arm1().dof()
        self._finish(marm, marm.arm2().dof(), marm.reverse_linact(), marm.torso_T_arm1(), marm.torso_T_arm2())


# ------------------------------------------------------------------------------------------------
# signed distance fields
# ------------------------------------------------------------------------------------------------
class _SdfBase:
    def __del__(self):
        try:
            _drop_handles(self, "sdf")
        except Exception:
            pass

    def _pack(self):
        d = _abi.SdfDesc()
        d.ndim = self.ndim
        d.rows, d.cols, d.nz = self._rows, self._cols, self._nz
        d.origin = (C.c_double * 3)(*self._origin)
        d.cell_size = self._cell
        d.data = _abi.dptr(self._wire)
        self.desc = d
        self._handles = {}

    def cell_size(self):
        return self._cell

    def x_count(self):
        return self._cols

    def y_count(self):
        return self._rows


class PlanarSDF(_SdfBase):
    """data: (rows, cols) matrix, data[r, c] = field at y index r, x index c (like the Eigen Matrix)."""
    ndim = 2

    def __init__(self, origin, cell_size, data):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim != 2:
            raise RuntimeError("[PlanarSDF] data must be a matrix")
        self._rows, self._cols, self._nz = data.shape[0], data.shape[1], 1
        o = np.asarray(origin, dtype=np.float64).ravel()
        self._origin = [o[0], o[1], 0.0]
        self._cell = float(cell_size)
        self._wire = np.ascontiguousarray(data.T)  # [col][row]
        self._pack()

    def origin(self):
        """PlanarSDF.h:123 (as an (x, y) array)."""
        return np.array(self._origin[:2])

    def signed_distance(self, r, c):
        """Raw field value data_(r, c) (PlanarSDF.h:119-121)."""
        return float(self._wire[c, r])


class SignedDistanceField(_SdfBase):
    """data: list of nz (rows, cols) matrices or an (nz, rows, cols) array."""
    ndim = 3

    def __init__(self, origin, cell_size, *args):
        o = np.asarray(origin, dtype=np.float64).ravel()
        self._origin = [o[0], o[1], o[2]]
        self._cell = float(cell_size)
        if len(args) == 1:
            data = np.asarray(args[0], dtype=np.float64)
            if data.ndim != 3:
                raise RuntimeError("[SignedDistanceField] data must be nz matrices of equal size")
            self._nz, self._rows, self._cols = data.shape
            self._wire = np.ascontiguousarray(data.transpose(0, 2, 1))  # [z][col][row]
            self._pack()
        elif len(args) == 3:
            self._rows, self._cols, self._nz = int(args[0]), int(args[1]), int(args[2])
            self._wire = np.zeros((self._nz, self._cols, self._rows), dtype=np.float64)
            self._pack()
        else:
            raise TypeError("SignedDistanceField(origin, cell_size, data) or (origin, cell_size, rows, cols, z)")

    def initFieldData(self, z_idx, field_layer):
        if z_idx >= self._nz:
            raise RuntimeError("[SignedDistanceField] matrix layer out of index")
        layer = np.asarray(field_layer, dtype=np.float64)
        self._wire[z_idx] = layer.T
        _drop_handles(self, "sdf")  # invalidate (and free) the device copies

    def z_count(self):
        return self._nz

    def origin(self):
        """SignedDistanceField.h:174 (as an (x, y, z) array)."""
        return np.array(self._origin)

    def signed_distance(self, r, c, z):
        """Raw field value data_[z](r, c) (SignedDistanceField.h:170-172)."""
        return float(self._wire[z, c, r])

    def saveSDF(self, filename):
        """SignedDistanceField::saveSDF (gpmp2/obstacle/SignedDistanceField.cpp:14-30): Boost archive chosen by the
        extension -- `.bin` binary, anything else text (`.xml` cannot be written by the reference either; see
        boost_archive.py, which also says why this format is parity-unpinned)."""
        from . import boost_archive
        boost_archive.save_sdf(filename, self._origin, self._rows, self._cols, self._nz, self._cell, self._wire)

    def loadSDF(self, filename):
        """SignedDistanceField::loadSDF (gpmp2/obstacle/SignedDistanceField.cpp:33-50): replaces this field."""
        from . import boost_archive
        o, self._rows, self._cols, self._nz, self._cell, wire = boost_archive.load_sdf(filename)
        self._origin = [float(o[0]), float(o[1]), float(o[2])]
        self._wire = np.ascontiguousarray(wire)
        _drop_handles(self, "sdf")
        self._pack()


def readSDFvolfile(filename_pre):
    """gpmp2::readSDFvolfile (gpmp2/utils/fileUtils.cpp:17-62): `<pre>.vol.head` holds `cols rows z`, the origin and the
    cell size as text; `<pre>.vol.data` the field values as text, x (column) outermost, then y (row), z innermost.
    Returns the SignedDistanceField, or None where the reference returns false (a file cannot be opened)."""
    try:
        with open(filename_pre + ".vol.head") as f:
            tok = f.read().split()
        cols, rows, nz = int(tok[0]), int(tok[1]), int(tok[2])
        origin = [float(tok[3]), float(tok[4]), float(tok[5])]
        res = float(tok[6])
        vals = np.fromfile(filename_pre + ".vol.data", dtype=np.float64, sep=" ", count=cols * rows * nz)
    except OSError:
        return None
    if vals.size != cols * rows * nz:
        raise RuntimeError("[readSDFvolfile] %s.vol.data holds %d values, header says %d" % (filename_pre, vals.size, cols * rows * nz))
    return SignedDistanceField(origin, res, vals.reshape(cols, rows, nz).transpose(2, 1, 0))   # -> [z][row][col]


def writeSDFvolfile(filename_pre, origin, cell_size, data):
    """Inverse of readSDFvolfile (the reference only reads this format): data = (nz, rows, cols) array."""
    data = np.asarray(data, dtype=np.float64)
    nz, rows, cols = data.shape
    with open(filename_pre + ".vol.head", "w") as f:
        f.write("%d %d %d\n%.17g %.17g %.17g\n%.17g\n" % (cols, rows, nz, origin[0], origin[1], origin[2], cell_size))
    with open(filename_pre + ".vol.data", "w") as f:
        np.savetxt(f, data.transpose(2, 1, 0).reshape(-1, nz), fmt="%.17g")


# ------------------------------------------------------------------------------------------------
# settings
# ------------------------------------------------------------------------------------------------
class TrajOptimizerSetting:
    GaussNewton, LM, Dogleg = _abi.OPT_GAUSS_NEWTON, _abi.OPT_LM, _abi.OPT_DOGLEG
    NONE, Error = 0, 1

    def __init__(self, system_dof):
        # defaults: gpmp2/planner/TrajOptimizerSetting.cpp:44-68
        D = int(system_dof)
        self.dof = D
        self.total_step = 10
        self.total_time = 1.0
        self.conf_prior_sigma = 0.0001
        self.vel_prior_sigma = 0.0001
        self.flag_pos_limit = False
        self.flag_vel_limit = False
        self.joint_pos_limits_up = 1e6 * np.ones(D)
        self.joint_pos_limits_down = -1e6 * np.ones(D)
        self.vel_limits = 1e6 * np.ones(D)
        self.pos_limit_thresh = 0.001 * np.ones(D)
        self.vel_limit_thresh = 0.001 * np.ones(D)
        self.pos_limit_sigma = 0.001 * np.ones(D)
        self.vel_limit_sigma = 0.001 * np.ones(D)
        self.epsilon = 0.2
        self.cost_sigma = 0.1
        self.obs_check_inter = 5
        self.Qc = np.eye(D)
        self.opt_type = self.Dogleg
        self.opt_verbosity = self.NONE
        self.final_iter_no_increase = True
        self.rel_thresh = 1e-2
        self.max_iter = 50
        # optional workspace goal on x_T (not part of the reference's struct: its hand-built graphs add the factor,
        # matlab/Arm3GoalReachExample.m:107) -- see set_workspace_goal
        self.goal_enabled = False
        self.goal_link = 0
        self.goal_keep_end_prior = False
        self.goal_sigma = 1.0
        self.goal_pos = np.zeros(3)
        self.goal_pos_batch = self.goal_R_batch = self.orient_R_batch = None   # per-problem targets (B rows) or None
        self.fix = None                                                        # fix_config_and_vel: pinned support state
        self.self_collision_data = None
        self.vehicle_dynamics_sigma = 0.0
        self.orient = None

    def set_vehicle_dynamics(self, sigma):
        """VehicleDynamicsFactorPose2Vector(x_i, v_i, sigma) on every support state of a Pose2MobileArm
        (gpmp2/dynamics/VehicleDynamicsFactorPose2Vector.h:46-79; matlab/MobileArm2FactorGraphExample.m:122-126).  0 = off."""
        self.vehicle_dynamics_sigma = float(sigma)

    def set_workspace_orientation(self, des_R, sigma, link=None, first_state=0, last_state=None):
        """GaussianPriorWorkspaceOrientationArm(x_i, arm, link, Rot3(des_R), Isotropic::Sigma(3, sigma)) on support states
        first_state..last_state (default: all) -- gpmp2/kinematics/GaussianPriorWorkspaceOrientation.h:40-72,
        matlab/WAMWorkspaceConstraintsExample.m:100-104.  des_R = None switches it off."""
        if des_R is None:
            self.orient = None
            return
        self.orient = dict(R=np.asarray(des_R, dtype=np.float64).reshape(3, 3).copy(), sigma=float(sigma),
                           link=-1 if link is None else int(link), first=int(first_state),
                           last=None if last_state is None else int(last_state))

    def set_workspace_goal(self, goal_point, sigma, link=None, keep_end_conf_prior=False):
        """GoalFactorArm(x_T, Isotropic::Sigma(3, sigma), arm, goal_point) (gpmp2/kinematics/GoalFactorArm.h:47-77;
        link = None -> the last joint frame) or GaussianPriorWorkspacePositionArm(x_T, arm, link, goal_point, ...)
        (GaussianPriorWorkspacePosition.h:46-76).  By default it replaces the end-configuration prior, as the
        reference's goal-reach example builds its graph."""
        self.goal_enabled = True
        self.goal_link = -1 if link is None else int(link)
        self.goal_keep_end_prior = bool(keep_end_conf_prior)
        self.goal_sigma = float(sigma)
        self.goal_pos = np.asarray(goal_point, dtype=np.float64).ravel().copy()

    def clear_workspace_goal(self):
        self.goal_enabled = False
        self.goal_pos_batch = self.goal_R_batch = None

    def set_workspace_goal_batch(self, goal_points=None, des_R=None):
        """One workspace target PER PROBLEM of the next call (the reference attaches the factor per graph: a batch of
        different queries has different goals).  goal_points (B, 3) replaces the shared goal point, des_R (B, 3, 3) the
        shared goal rotation of a pose goal; None = keep the shared value.  Call set_workspace_goal /
        set_workspace_pose_goal first (they hold sigma, link and the shared fallbacks)."""
        self.goal_pos_batch = None if goal_points is None else np.ascontiguousarray(np.asarray(goal_points, dtype=np.float64).reshape(-1, 3))
        self.goal_R_batch = None if des_R is None else np.ascontiguousarray(np.asarray(des_R, dtype=np.float64).reshape(-1, 9))

    def fix_config_and_vel(self, state_idx, conf_fix, vel_fix):
        """ISAM2TrajOptimizer::fixConfigAndVel (gpmp2/planner/ISAM2TrajOptimizer-inl.h:160-168) for a batch of replanning
        problems in lockstep: PriorFactor(x_k, conf_fix[p], conf_prior_model) + PriorFactor(v_k, vel_fix[p], vel_prior_model)
        on support state k = state_idx of every problem p; conf_fix, vel_fix: (B, dof).  Together with new end_conf /
        end_vel rows (changeGoalConfigAndVel) and init_traj = the previous result (initValues) this is the warm-started
        re-solve of the replanner; clear_fixed_state() removes it."""
        c = np.ascontiguousarray(np.asarray(conf_fix, dtype=np.float64).reshape(-1, self.dof))
        v = np.ascontiguousarray(np.asarray(vel_fix, dtype=np.float64).reshape(-1, self.dof))
        if c.shape != v.shape:
            raise RuntimeError("[TrajOptimizerSetting] ERROR: conf_fix and vel_fix have different shapes.")
        self.fix = {"index": int(state_idx), "conf": c, "vel": v}

    def clear_fixed_state(self):
        self.fix = None

    def set_workspace_orientation_batch(self, des_R=None):
        """One desired rotation per problem, (B, 3, 3), for the orientation priors of set_workspace_orientation."""
        self.orient_R_batch = None if des_R is None else np.ascontiguousarray(np.asarray(des_R, dtype=np.float64).reshape(-1, 9))

    def set_workspace_pose_goal(self, des_R, des_t, sigma, link=None, keep_end_conf_prior=False):
        """GaussianPriorWorkspacePoseArm(x_T, arm, link, Pose3(Rot3(des_R), des_t), Isotropic::Sigma(6, sigma))
        (gpmp2/kinematics/GaussianPriorWorkspacePose.h:40-70) -- the end-state factor of
        matlab/WAMWorkspaceConstraintsExample.m:94-96; replaces the end-configuration prior unless keep_end_conf_prior."""
        self.set_workspace_goal(des_t, sigma, link, keep_end_conf_prior)
        self.goal_enabled = 2
        self.goal_R = np.asarray(des_R, dtype=np.float64).reshape(3, 3).copy()

    def set_self_collision(self, data):
        """SelfCollisionArm(x_i, arm, data) on every support state (gpmp2/obstacle/SelfCollision.h:38-60): rows of
        (sphere A id, sphere B id, epsilon, sigma).  None / empty = off."""
        self.self_collision_data = None if data is None or len(data) == 0 else \
            np.ascontiguousarray(np.asarray(data, dtype=np.float64).reshape(-1, 4))

    # setters, same names as the reference (TrajOptimizerSetting.h:61-99)
    def set_total_step(self, step): self.total_step = int(step)
    def set_total_time(self, time): self.total_time = float(time)
    def set_conf_prior_model(self, sigma): self.conf_prior_sigma = float(sigma)
    def set_vel_prior_model(self, sigma): self.vel_prior_sigma = float(sigma)
    def set_flag_pos_limit(self, flag): self.flag_pos_limit = bool(flag)
    def set_flag_vel_limit(self, flag): self.flag_vel_limit = bool(flag)
    def set_joint_pos_limits_up(self, v): self.joint_pos_limits_up = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_joint_pos_limits_down(self, v): self.joint_pos_limits_down = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_vel_limits(self, v): self.vel_limits = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_pos_limit_thresh(self, v): self.pos_limit_thresh = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_vel_limit_thresh(self, v): self.vel_limit_thresh = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_pos_limit_model(self, v): self.pos_limit_sigma = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_vel_limit_model(self, v): self.vel_limit_sigma = np.asarray(v, dtype=np.float64).ravel().copy()
    def set_epsilon(self, eps): self.epsilon = float(eps)
    def set_cost_sigma(self, sigma): self.cost_sigma = float(sigma)
    def set_obs_check_inter(self, inter): self.obs_check_inter = int(inter)
    def set_Qc_model(self, Qc): self.Qc = np.asarray(Qc, dtype=np.float64).reshape(self.dof, self.dof).copy()
    def setGaussNewton(self): self.opt_type = self.GaussNewton
    def setLM(self): self.opt_type = self.LM
    def setDogleg(self): self.opt_type = self.Dogleg
    def set_rel_thresh(self, thresh): self.rel_thresh = float(thresh)
    def set_max_iter(self, it): self.max_iter = int(it)
    def setVerbosityNone(self): self.opt_verbosity = self.NONE
    def setVerbosityError(self): self.opt_verbosity = self.Error
    def setOptimizationNoIncrase(self, flag): self.final_iter_no_increase = bool(flag)

    def pack(self):
        """-> (gpmp2b_setting, keepalive list)"""
        D = self.dof
        keep = []

        def vec(v, name):
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
            if a.size != D:
                raise RuntimeError("[TrajOptimizerSetting] ERROR: %s dim does not fit." % name)
            keep.append(a)
            return _abi.dptr(a)

        s = _abi.Setting()
        s.dof, s.total_step, s.total_time = D, self.total_step, self.total_time
        s.conf_prior_sigma, s.vel_prior_sigma = self.conf_prior_sigma, self.vel_prior_sigma
        s.flag_pos_limit, s.flag_vel_limit = int(self.flag_pos_limit), int(self.flag_vel_limit)
        s.joint_pos_limits_up = vec(self.joint_pos_limits_up, "joint_pos_limits_up")
        s.joint_pos_limits_down = vec(self.joint_pos_limits_down, "joint_pos_limits_down")
        s.vel_limits = vec(self.vel_limits, "vel_limits")
        s.pos_limit_thresh = vec(self.pos_limit_thresh, "pos_limit_thresh")
        s.vel_limit_thresh = vec(self.vel_limit_thresh, "vel_limit_thresh")
        s.pos_limit_sigma = vec(self.pos_limit_sigma, "pos_limit_model")
        s.vel_limit_sigma = vec(self.vel_limit_sigma, "vel_limit_model")
        s.epsilon, s.cost_sigma = self.epsilon, self.cost_sigma
        s.obs_check_inter, s.opt_type = self.obs_check_inter, self.opt_type
        q = np.ascontiguousarray(np.asarray(self.Qc, dtype=np.float64).reshape(D, D))
        keep.append(q)
        s.Qc = _abi.dptr(q)
        s.opt_verbosity = self.opt_verbosity
        s.final_iter_no_increase = int(self.final_iter_no_increase)
        s.rel_thresh, s.max_iter = self.rel_thresh, self.max_iter
        if self.goal_enabled:
            if self.goal_pos.size != 3:
                raise RuntimeError("[TrajOptimizerSetting] ERROR: workspace goal must be a 3-vector.")
            s.goal_enabled = int(self.goal_enabled)       # 1 position goal, 2 pose goal
            if int(self.goal_enabled) == 2:
                for k in range(9):
                    s.goal_R[k] = float(self.goal_R.ravel()[k])
            s.goal_link = self.goal_link      # -1 = last joint frame (resolved by the library against the robot)
            s.goal_keep_end_prior = int(self.goal_keep_end_prior)
            s.goal_sigma = self.goal_sigma
            for k in range(3):
                s.goal_pos[k] = float(self.goal_pos[k])
        if self.self_collision_data is not None:
            keep.append(self.self_collision_data)
            s.n_self_collision = self.self_collision_data.shape[0]
            s.self_collision_data = _abi.dptr(self.self_collision_data)
        s.vehicle_dynamics_sigma = self.vehicle_dynamics_sigma
        if self.orient is not None:
            o = self.orient
            s.orient_enabled, s.orient_link, s.orient_sigma = 1, o["link"], o["sigma"]
            s.orient_state_first = o["first"]
            s.orient_state_last = self.total_step if o["last"] is None else o["last"]
            for k in range(9):
                s.orient_R[k] = float(o["R"].ravel()[k])
            if self.orient_R_batch is not None:
                keep.append(self.orient_R_batch)
                s.orient_R_batch = self.orient_R_batch.ctypes.data
        if self.goal_enabled:
            if self.goal_pos_batch is not None:
                keep.append(self.goal_pos_batch)
                s.goal_pos_batch = self.goal_pos_batch.ctypes.data
            if self.goal_R_batch is not None and self.goal_enabled == 2:
                keep.append(self.goal_R_batch)
                s.goal_R_batch = self.goal_R_batch.ctypes.data
        if getattr(self, "fix", None) is not None:
            keep += [self.fix["conf"], self.fix["vel"]]
            s.fix_enabled, s.fix_state_index = 1, self.fix["index"]
            s.fix_conf, s.fix_vel = self.fix["conf"].ctypes.data, self.fix["vel"].ctypes.data
        return s, keep

    def batch_rows(self):
        """Number of rows of the per-problem target arrays (None if there are none): the call checks it against B."""
        rows = [a.shape[0] for a in (self.goal_pos_batch if self.goal_enabled else None,
                                     self.goal_R_batch if self.goal_enabled == 2 else None,
                                     self.orient_R_batch if self.orient is not None else None,
                                     self.fix["conf"] if getattr(self, "fix", None) is not None else None) if a is not None]
        if rows and min(rows) != max(rows):
            raise RuntimeError("[TrajOptimizerSetting] ERROR: per-problem target arrays have different lengths.")
        return rows[0] if rows else None


# ------------------------------------------------------------------------------------------------
# context + batched planner
# ------------------------------------------------------------------------------------------------
class Context:
    """One gpmp2b_ctx (one per thread and device)."""

    _serial = itertools.count(1)   # unique token per Context: id() can be reused after a Context is collected
    _live = weakref.WeakValueDictionary()

    def __init__(self, device=0):
        self.token = next(Context._serial)
        Context._live[self.token] = self
        self._owners = weakref.WeakSet()   # models / fields holding a device handle of this context
        self.lib = _abi.load_library()
        h = C.c_void_p()
        rc = self.lib.gpmp2b_create(int(device), C.byref(h))
        if rc != _abi.OK:
            raise RuntimeError("gpmp2b_create(device=%d) failed with status %d: no usable CUDA device "
                               "(there is no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != _abi.OK:
            msg = self.lib.gpmp2b_last_error(self.h).decode()
            if rc == _abi.ERR_INVALID_ARG:
                raise RuntimeError(msg)  # the reference throws std::runtime_error
            raise RuntimeError("gpmp2b status %d: %s" % (rc, msg))

    def robot_handle(self, model):
        key = self.token
        if key not in model._handles:
            h = C.c_void_p()
            self.check(self.lib.gpmp2b_robot_upload(self.h, C.byref(model.desc), C.byref(h)))
            model._handles[key] = h
            self._owners.add(model)
        return model._handles[key]

    def sdf_handle(self, sdf):
        key = self.token
        if key not in sdf._handles:
            h = C.c_void_p()
            self.check(self.lib.gpmp2b_sdf_upload(self.h, C.byref(sdf.desc), C.byref(h)))
            sdf._handles[key] = h
            self._owners.add(sdf)
        return sdf._handles[key]

    def release_handle(self, kind, h):
        """Free one device handle of this context (kind: "robot" | "sdf")."""
        if self.h:
            (self.lib.gpmp2b_sdf_free if kind == "sdf" else self.lib.gpmp2b_robot_free)(self.h, h)

    def launch_count(self):
        return int(self.lib.gpmp2b_launch_count(self.h))

    def last_kernel_stats(self):
        ms = C.c_double()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self.lib.gpmp2b_last_kernel_stats(self.h, C.byref(ms), C.byref(a), C.byref(b), C.byref(c)))
        return {"kernel_ms": ms.value, "linearizations": a.value, "solves": b.value, "error_evals": c.value}

    def measure_peaks(self):
        out = np.zeros(3)
        self.check(self.lib.gpmp2b_measure_peaks(self.h, _abi.dptr(out)))
        return {"fp64_tflops": out[0], "l2_gather_useful_gbs": out[1], "l2_gather_sector_gbs": out[2]}

    def close(self):
        if self.h:
            for o in list(self._owners):          # the handles die with the ctx: drop the cached copies
                o._handles.pop(self.token, None)
            self._owners.clear()
            self.lib.gpmp2b_destroy(self.h)
            self.h = None
        Context._live.pop(self.token, None)


def _drop_handles(obj, kind):
    """Invalidate the cached device copies of a model / field: free them in their (still open) contexts."""
    for token, h in list(getattr(obj, "_handles", {}).items()):
        ctx = Context._live.get(token)
        if ctx is not None:
            ctx.release_handle(kind, h)
    obj._handles = {}


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _as2d(a, B, n, name):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape != (B, n):
        raise RuntimeError("%s has shape %s, expected (%d, %d)" % (name, a.shape, B, n))
    return a


def batch_optimize(model, sdf, start_conf, start_vel, end_conf, end_vel, init_traj, setting, ctx=None):
    """B independent problems through gpmp2b_batch_optimize with HOST buffers.

    start_conf/end_conf/start_vel/end_vel: (B, D); init_traj: (B, 2*N*D) in the wire layout
    [x_0..x_T | v_0..v_T], or None for the straight line from start_conf to end_conf built on the device.
    Returns dict(traj, error, coll_cost, iters, status).
    """
    ctx = ctx or default_context()
    D = setting.dof
    if model.dof() != D:
        raise RuntimeError("setting.dof != robot dof")
    N = setting.total_step + 1
    if init_traj is None:     # straight-line initialisation on the device (TrajUtils.cpp:23-73), no trajectory upload
        B = np.asarray(start_conf).reshape(-1, D).shape[0]
    else:
        init_traj = np.ascontiguousarray(np.asarray(init_traj, dtype=np.float64))
        if init_traj.ndim == 1:
            init_traj = init_traj.reshape(1, -1)
        B = init_traj.shape[0]
        init_traj = _as2d(init_traj, B, 2 * N * D, "init_traj")
    sc, sv = _as2d(start_conf, B, D, "start_conf"), _as2d(start_vel, B, D, "start_vel")
    ec, ev = _as2d(end_conf, B, D, "end_conf"), _as2d(end_vel, B, D, "end_vel")
    out = np.empty((B, 2 * N * D))
    err, cc = np.empty(B), np.empty(B)
    iters, status = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    if setting.batch_rows() not in (None, B):
        raise RuntimeError("per-problem workspace targets: %d rows for %d problems" % (setting.batch_rows(), B))
    s, keep = setting.pack()
    ctx.check(ctx.lib.gpmp2b_batch_optimize(
        ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(s), B,
        sc.ctypes.data, sv.ctypes.data, ec.ctypes.data, ev.ctypes.data,
        None if init_traj is None else init_traj.ctypes.data, out.ctypes.data,
        err.ctypes.data, cc.ctypes.data, iters.ctypes.data, status.ctypes.data, _abi.MEM_HOST, None))
    del keep
    return {"traj": out, "error": err, "coll_cost": cc, "iters": iters, "status": status}


def batch_optimize_multi(contexts, model, sdf, start_conf, start_vel, end_conf, end_vel, init_traj, setting):
    """gpmp2b_batch_optimize_multi with HOST buffers: the batch is cut into len(contexts) contiguous shards, one per
    context (one Context per GPU; two on the same GPU are allowed), each driven by its own host thread inside the
    library.  Same arguments and result as batch_optimize."""
    n = len(contexts)
    D = setting.dof
    if model.dof() != D:
        raise RuntimeError("setting.dof != robot dof")
    N = setting.total_step + 1
    if init_traj is None:
        B = np.asarray(start_conf).reshape(-1, D).shape[0]
    else:
        init_traj = np.ascontiguousarray(np.asarray(init_traj, dtype=np.float64))
        if init_traj.ndim == 1:
            init_traj = init_traj.reshape(1, -1)
        B = init_traj.shape[0]
        init_traj = _as2d(init_traj, B, 2 * N * D, "init_traj")
    sc, sv = _as2d(start_conf, B, D, "start_conf"), _as2d(start_vel, B, D, "start_vel")
    ec, ev = _as2d(end_conf, B, D, "end_conf"), _as2d(end_vel, B, D, "end_vel")
    out = np.empty((B, 2 * N * D))
    err, cc = np.empty(B), np.empty(B)
    iters, status = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    s, keep = setting.pack()
    VP = C.c_void_p * n
    hs = VP(*[c.h for c in contexts])
    rs = VP(*[c.robot_handle(model) for c in contexts])
    fs = VP(*[c.sdf_handle(sdf) for c in contexts])
    rc = contexts[0].lib.gpmp2b_batch_optimize_multi(
        n, hs, rs, fs, C.byref(s), B, sc.ctypes.data, sv.ctypes.data, ec.ctypes.data, ev.ctypes.data,
        None if init_traj is None else init_traj.ctypes.data, out.ctypes.data,
        err.ctypes.data, cc.ctypes.data, iters.ctypes.data, status.ctypes.data, _abi.MEM_HOST)
    if rc != 0:
        for c in contexts:
            msg = c.lib.gpmp2b_last_error(c.h)
            if msg:
                raise RuntimeError("gpmp2b_batch_optimize_multi failed (%d): %s" % (rc, msg.decode()))
        raise RuntimeError("gpmp2b_batch_optimize_multi failed (%d)" % rc)
    del keep
    return {"traj": out, "error": err, "coll_cost": cc, "iters": iters, "status": status}


def batch_linearize(model, sdf, start_conf, start_vel, end_conf, end_vel, traj, setting, ctx=None):
    """Debug/parity entry (gpmp2b_linearize): block-tridiagonal H, g and graph error at `traj`."""
    ctx = ctx or default_context()
    D, N = setting.dof, setting.total_step + 1
    b = 2 * D
    traj = np.ascontiguousarray(np.asarray(traj, dtype=np.float64))
    if traj.ndim == 1:
        traj = traj.reshape(1, -1)
    B = traj.shape[0]
    traj = _as2d(traj, B, 2 * N * D, "traj")
    sc, sv = _as2d(start_conf, B, D, "start_conf"), _as2d(start_vel, B, D, "start_vel")
    ec, ev = _as2d(end_conf, B, D, "end_conf"), _as2d(end_vel, B, D, "end_vel")
    Hd, Ho = np.empty((B, N, b, b)), np.empty((B, N - 1, b, b))
    g, err = np.empty((B, N, b)), np.empty(B)
    s, keep = setting.pack()
    ctx.check(ctx.lib.gpmp2b_linearize(
        ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(s), B,
        sc.ctypes.data, sv.ctypes.data, ec.ctypes.data, ev.ctypes.data, traj.ctypes.data,
        Hd.ctypes.data, Ho.ctypes.data, g.ctypes.data, err.ctypes.data, _abi.MEM_HOST, None))
    del keep
    return {"Hdiag": Hd, "Hoff": Ho, "g": g, "error": err}


def batch_obstacle_errors(model, sdf, traj, setting, want_centers=True, ctx=None):
    """Debug/parity entry (gpmp2b_obstacle_errors)."""
    ctx = ctx or default_context()
    D, N, K = setting.dof, setting.total_step + 1, setting.obs_check_inter
    S = model.nr_body_spheres()
    Cn = N + (N - 1) * K
    traj = np.ascontiguousarray(np.asarray(traj, dtype=np.float64))
    if traj.ndim == 1:
        traj = traj.reshape(1, -1)
    B = traj.shape[0]
    traj = _as2d(traj, B, 2 * N * D, "traj")
    err = np.empty((B, Cn, S))
    ctr = np.empty((B, Cn, S, 3)) if want_centers else None
    s, keep = setting.pack()
    ctx.check(ctx.lib.gpmp2b_obstacle_errors(
        ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(s), B, traj.ctypes.data,
        err.ctypes.data, ctr.ctypes.data if want_centers else None, _abi.MEM_HOST, None))
    del keep
    return {"err": err, "centers": ctr}


def batch_collision_cost(model, sdf, traj, setting, ctx=None):
    ctx = ctx or default_context()
    D, N = setting.dof, setting.total_step + 1
    traj = np.ascontiguousarray(np.asarray(traj, dtype=np.float64))
    if traj.ndim == 1:
        traj = traj.reshape(1, -1)
    B = traj.shape[0]
    traj = _as2d(traj, B, 2 * N * D, "traj")
    out = np.empty(B)
    s, keep = setting.pack()
    ctx.check(ctx.lib.gpmp2b_collision_cost(
        ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(s), B, traj.ctypes.data,
        out.ctypes.data, _abi.MEM_HOST, None))
    del keep
    return out


# ------------------------------------------------------------------------------------------------
# Values <-> wire layout
# ------------------------------------------------------------------------------------------------
def values_to_traj(values, total_step, D):
    N = total_step + 1
    t = np.empty(2 * N * D)
    for i in range(N):
        x = values.atVector(symbol('x', i))
        x = x.flat() if isinstance(x, Pose2Vector) else np.asarray(x, dtype=np.float64).ravel()
        v = np.asarray(values.atVector(symbol('v', i)), dtype=np.float64).ravel()
        if x.size != D or v.size != D:
            raise RuntimeError("init_values: x(%d)/v(%d) dimension does not fit dof" % (i, i))
        t[i * D:(i + 1) * D] = x
        t[(N + i) * D:(N + i + 1) * D] = v
    return t


def traj_to_values(traj, total_step, D, lie=False):
    N = total_step + 1
    vals = Values()
    for i in range(N):
        x = traj[i * D:(i + 1) * D].copy()
        vals.insert(symbol('x', i), Pose2Vector.from_flat(x) if lie else x)
        vals.insert(symbol('v', i), traj[(N + i) * D:(N + i + 1) * D].copy())
    return vals


def _flat_state(x):
    return x.flat() if isinstance(x, Pose2Vector) else np.asarray(x, dtype=np.float64).ravel()


_DUMMY_FIELD = []


def _dummy_field():
    """A 2 x 2 x 2 zero field for device entry points that need an SDF argument but whose SDF-dependent output is ignored."""
    if not _DUMMY_FIELD:
        _DUMMY_FIELD.append(SignedDistanceField([0.0, 0.0, 0.0], 1.0, np.zeros((2, 2, 2))))
    return _DUMMY_FIELD[0]


def _single(model, sdf, ndim, start_conf, start_vel, end_conf, end_vel, init_values, setting, lie):
    if sdf.ndim != ndim:
        raise TypeError("wrong SDF type for this planner")
    D = setting.dof
    t0 = values_to_traj(init_values, setting.total_step, D)
    r = batch_optimize(model, sdf, _flat_state(start_conf), np.asarray(start_vel, dtype=np.float64),
                       _flat_state(end_conf), np.asarray(end_vel, dtype=np.float64), t0, setting)
    return traj_to_values(r["traj"][0], setting.total_step, D, lie)


# reference signatures (B = 1), gpmp2/planner/BatchTrajOptimizer.h:43-104
def BatchTrajOptimize2DArm(arm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(arm, sdf, 2, start_conf, start_vel, end_conf, end_vel, init_values, setting, False)


def BatchTrajOptimize3DArm(arm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(arm, sdf, 3, start_conf, start_vel, end_conf, end_vel, init_values, setting, False)


def BatchTrajOptimizePose2MobileArm2D(marm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(marm, sdf, 2, start_conf, start_vel, end_conf, end_vel, init_values, setting, True)


def BatchTrajOptimizePose2MobileArm(marm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(marm, sdf, 3, start_conf, start_vel, end_conf, end_vel, init_values, setting, True)


# the other Pose2Vector robots of gpmp2/planner/BatchTrajOptimizer.cpp:92-128 (3-D fields, as in the reference)
def BatchTrajOptimizePose2Mobile2Arms(marm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(marm, sdf, 3, start_conf, start_vel, end_conf, end_vel, init_values, setting, True)


def BatchTrajOptimizePose2MobileVetLinArm(marm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(marm, sdf, 3, start_conf, start_vel, end_conf, end_vel, init_values, setting, True)


def BatchTrajOptimizePose2MobileVetLin2Arms(marm, sdf, start_conf, start_vel, end_conf, end_vel, init_values, setting):
    return _single(marm, sdf, 3, start_conf, start_vel, end_conf, end_vel, init_values, setting, True)


def _coll(model, sdf, ndim, result, setting):
    if sdf.ndim != ndim:
        raise TypeError("wrong SDF type for this function")
    # like the reference, iterate i < result.size()/2 (BatchTrajOptimizer-inl.h:96)
    steps = result.size() // 2 - 1
    t = values_to_traj(result, steps, setting.dof)
    import copy
    st = copy.copy(setting)
    st.total_step = steps
    return float(batch_collision_cost(model, sdf, t, st)[0])


def CollisionCost2DArm(arm, sdf, result, setting): return _coll(arm, sdf, 2, result, setting)
def CollisionCost3DArm(arm, sdf, result, setting): return _coll(arm, sdf, 3, result, setting)
def CollisionCostPose2MobileArm2D(marm, sdf, result, setting): return _coll(marm, sdf, 2, result, setting)
def CollisionCostPose2MobileArm(marm, sdf, result, setting): return _coll(marm, sdf, 3, result, setting)
def CollisionCostPose2Mobile2Arms(marm, sdf, result, setting): return _coll(marm, sdf, 3, result, setting)
def CollisionCostPose2MobileVetLinArm(marm, sdf, result, setting): return _coll(marm, sdf, 3, result, setting)
def CollisionCostPose2MobileVetLin2Arms(marm, sdf, result, setting): return _coll(marm, sdf, 3, result, setting)


# gpmp2/planner/TrajUtils.cpp:25-50 -- host-side input helper (produces init_values)
def initArmTrajStraightLine(init_conf, end_conf, total_step):
    init_conf = np.asarray(init_conf, dtype=np.float64).ravel()
    end_conf = np.asarray(end_conf, dtype=np.float64).ravel()
    vals = Values()
    for i in range(total_step + 1):
        if i == 0:
            conf = init_conf.copy()
        elif i == total_step:
            conf = end_conf.copy()
        else:
            conf = float(i) / float(total_step) * end_conf + (1.0 - float(i) / float(total_step)) * init_conf
        vals.insert(symbol('x', i), conf)
    avg_vel = (end_conf - init_conf) / float(total_step)
    for i in range(total_step + 1):
        vals.insert(symbol('v', i), avg_vel.copy())
    return vals


def straight_line_traj(start_conf, end_conf, total_step):
    """Batched initArmTrajStraightLine in the wire layout: (B, D) x (B, D) -> (B, 2*N*D)."""
    s = np.asarray(start_conf, dtype=np.float64)
    e = np.asarray(end_conf, dtype=np.float64)
    B, D = s.shape
    N = total_step + 1
    t = np.empty((B, 2 * N, D))
    for i in range(N):
        if i == 0:
            t[:, i] = s
        elif i == total_step:
            t[:, i] = e
        else:
            r = float(i) / float(total_step)
            t[:, i] = r * e + (1.0 - r) * s
    t[:, N:] = ((e - s) / float(total_step))[:, None, :]
    return t.reshape(B, 2 * N * D)


def batch_optimize_device(model, sdf, setting, B, start_conf, start_vel, end_conf, end_vel, init_traj, out_traj,
                          out_error=0, out_coll_cost=0, out_iters=0, out_status=0, stream=0, ctx=None):
    """gpmp2b_batch_optimize with DEVICE pointers (ints, e.g. torch.Tensor.data_ptr()): stream-ordered on
    `stream` (a cudaStream_t as int, 0 = default stream), asynchronous, zero-copy."""
    ctx = ctx or default_context()
    s, keep = setting.pack()
    vp = lambda p: C.c_void_p(int(p)) if p else None  # noqa: E731
    ctx.check(ctx.lib.gpmp2b_batch_optimize(
        ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(s), int(B),
        vp(start_conf), vp(start_vel), vp(end_conf), vp(end_vel), vp(init_traj), vp(out_traj),
        vp(out_error), vp(out_coll_cost), vp(out_iters), vp(out_status), _abi.MEM_DEVICE, vp(stream)))
    del keep


# ------------------------------------------------------------------------------------------------
# trajectory utilities on the device (gpmp2/planner/TrajUtils.cpp)
# ------------------------------------------------------------------------------------------------
def batch_init_straight_line(start_conf, end_conf, total_step, lie=False, ctx=None):
    """initArmTrajStraightLine (lie=False) / initPose2VectorTrajStraightLine (lie=True, states (x, y, theta, q...))
    for B (start, end) pairs on the device: (B, D) x (B, D) -> (B, 2*N*D) in the wire layout."""
    ctx = ctx or default_context()
    s = np.ascontiguousarray(np.asarray(start_conf, dtype=np.float64))
    e = np.ascontiguousarray(np.asarray(end_conf, dtype=np.float64))
    if s.ndim == 1:
        s, e = s.reshape(1, -1), e.reshape(1, -1)
    if s.shape != e.shape:
        raise RuntimeError("start/end configuration dimensions do not match")
    B, D = s.shape
    out = np.empty((B, 2 * (total_step + 1) * D))
    ctx.check(ctx.lib.gpmp2b_init_straight_line(
        ctx.h, _abi.ROBOT_POSE2_MOBILE_ARM if lie else _abi.ROBOT_ARM, D, int(total_step), B,
        s.ctypes.data, e.ctypes.data, out.ctypes.data, _abi.MEM_HOST, None))
    return out


def batch_interpolate_traj(traj, dof, total_step, delta_t, inter_step, Qc=None, lie=False, start_index=0,
                           end_index=None, ctx=None):
    """interpolateArmTraj / interpolatePose2MobileArmTraj for B trajectories on the device:
    (B, 2*N*D) -> (B, 2*Nout*D), Nout = (end_index - start_index) * (inter_step + 1) + 1."""
    ctx = ctx or default_context()
    end_index = total_step if end_index is None else int(end_index)
    t = np.ascontiguousarray(np.asarray(traj, dtype=np.float64))
    if t.ndim == 1:
        t = t.reshape(1, -1)
    B = t.shape[0]
    t = _as2d(t, B, 2 * (total_step + 1) * dof, "traj")
    if not (0 <= start_index < end_index <= total_step):
        raise RuntimeError("interpolate: bad start_index / end_index")
    nout = (end_index - start_index) * (inter_step + 1) + 1
    out = np.empty((B, 2 * nout * dof))
    q = None if Qc is None else np.ascontiguousarray(np.asarray(Qc, dtype=np.float64).reshape(dof, dof))
    ctx.check(ctx.lib.gpmp2b_interpolate_traj(
        ctx.h, _abi.ROBOT_POSE2_MOBILE_ARM if lie else _abi.ROBOT_ARM, dof, int(total_step), float(delta_t),
        None if q is None else q.ctypes.data, int(inter_step), int(start_index), end_index, B,
        t.ctypes.data, out.ctypes.data, _abi.MEM_HOST, None))
    return out


def select_best(error, coll_cost=None, restarts=1, coll_tol=0.0, ctx=None):
    """Best of `restarts` consecutive problems per query: (best index into the batch [G], feasible flag [G]).
    Feasible = collision cost <= coll_tol (CollisionCost* of the result, optionally of its densification)."""
    ctx = ctx or default_context()
    e = np.ascontiguousarray(np.asarray(error, dtype=np.float64).ravel())
    if restarts < 1 or e.size % restarts:
        raise RuntimeError("select_best: batch size is not a multiple of restarts")
    G = e.size // restarts
    c = None if coll_cost is None else np.ascontiguousarray(np.asarray(coll_cost, dtype=np.float64).ravel())
    if c is not None and c.size != e.size:
        raise RuntimeError("select_best: error / collision cost sizes differ")
    best = np.empty(G, dtype=np.int64)
    feas = np.empty(G, dtype=np.int32)
    ctx.check(ctx.lib.gpmp2b_select_best(ctx.h, G, int(restarts), e.ctypes.data, None if c is None else c.ctypes.data,
                                         float(coll_tol), best.ctypes.data, feas.ctypes.data, _abi.MEM_HOST, None))
    return best, feas


def _values_steps(values):
    n = 0
    while symbol('x', n) in values:
        n += 1
    if n < 2:
        raise RuntimeError("values need at least x0, x1")
    return n - 1


def initPose2VectorTrajStraightLine(init_pose, init_conf, end_pose, end_conf, total_step, ctx=None):
    """gpmp2/planner/TrajUtils.cpp:51-73 (device)."""
    s = Pose2Vector(init_pose, init_conf).flat()
    e = Pose2Vector(end_pose, end_conf).flat()
    t = batch_init_straight_line(s, e, total_step, lie=True, ctx=ctx)[0]
    return traj_to_values(t, total_step, s.size, lie=True)


def initPose2TrajStraightLine(init_pose, end_pose, total_step, ctx=None):
    """gpmp2/planner/TrajUtils.cpp:76-93 (device): a bare Pose2 trajectory is the Pose2Vector one with an empty arm (dof 3);
    the values under x(i) are Pose2."""
    vals = initPose2VectorTrajStraightLine(init_pose, np.zeros(0), end_pose, np.zeros(0), total_step, ctx=ctx)
    return _pose2vector_values_to_pose2(vals)


def _pose2vector_values_to_pose2(vals):
    out = Values()
    for k, v in vals.items():
        out.insert(k, v.pose() if isinstance(v, Pose2Vector) else v)
    return out


def interpolatePose2Traj(values, Qc_model, delta_t, inter_step, start_index, end_index, ctx=None):
    """gpmp2/planner/TrajUtils.cpp:239-275 (device): GaussianProcessInterpolatorPose2 = the Lie interpolator on SE(2), i.e.
    the Pose2Vector one at dof 3."""
    v3 = Values()
    for k, v in values.items():
        v3.insert(k, Pose2Vector(v, np.zeros(0)) if isinstance(v, Pose2) else v)
    return _pose2vector_values_to_pose2(_interpolate_values(v3, Qc_model, delta_t, inter_step, start_index, end_index, True, ctx))


def _interpolate_values(values, Qc_model, delta_t, inter_step, start_index, end_index, lie, ctx):
    total_step = _values_steps(values)
    x0 = values.atVector(symbol('x', 0))
    D = x0.flat().size if isinstance(x0, Pose2Vector) else np.asarray(x0).size
    start_index = 0 if start_index is None else int(start_index)
    end_index = total_step if end_index is None else int(end_index)
    t = values_to_traj(values, total_step, D)
    out = batch_interpolate_traj(t, D, total_step, delta_t, inter_step, Qc=Qc_model, lie=lie, start_index=start_index,
                                 end_index=end_index, ctx=ctx)[0]
    return traj_to_values(out, (end_index - start_index) * (inter_step + 1), D, lie=lie)


def interpolateArmTraj(values, Qc_model, delta_t, inter_step, start_index=None, end_index=None, ctx=None):
    """gpmp2/planner/TrajUtils.cpp:96-196 (both overloads; device).  Qc_model: the Qc covariance matrix."""
    return _interpolate_values(values, Qc_model, delta_t, inter_step, start_index, end_index, False, ctx)


def interpolatePose2MobileArmTraj(values, Qc_model, delta_t, inter_step, start_index, end_index, ctx=None):
    """gpmp2/planner/TrajUtils.cpp:199-237 (device)."""
    return _interpolate_values(values, Qc_model, delta_t, inter_step, start_index, end_index, True, ctx)


# ------------------------------------------------------------------------------------------------
# signed distance fields from occupancy grids on the device (matlab/+gpmp2/signedDistanceField{2D,3D}.m)
# ------------------------------------------------------------------------------------------------
def _sdf_from_occupancy(ground_truth_map, cell_size, ndim, single_precision, ctx):
    ctx = ctx or default_context()
    m = np.ascontiguousarray(np.asarray(ground_truth_map, dtype=np.float64))
    if m.ndim != ndim:
        raise RuntimeError("occupancy map must be %d-dimensional" % ndim)
    d = _abi.SdfDesc()
    d.ndim = ndim
    # the transform is isotropic: any axis order works as long as input and output agree; C order (n0, n1, n2) is
    # read as [z][col][row]
    if ndim == 3:
        d.nz, d.cols, d.rows = m.shape
    else:
        d.nz, (d.cols, d.rows) = 1, m.shape
    d.cell_size = float(cell_size)
    d.origin[0] = d.origin[1] = d.origin[2] = 0.0
    d.data = m.ctypes.data_as(_abi.c_double_p)
    out = np.empty_like(m)
    ctx.check(ctx.lib.gpmp2b_sdf_from_occupancy(ctx.h, C.byref(d), int(bool(single_precision)), None, out.ctypes.data))
    return out


def signedDistanceField3D(ground_truth_map, cell_size, single_precision=True, ctx=None):
    """matlab/+gpmp2/signedDistanceField3D.m:16-33 on the device: occupancy > 0.75 is an obstacle; returns
    (distance to the nearest obstacle cell - distance to the nearest free cell) * cell_size, same shape as the map.
    single_precision=True reproduces MATLAB's bwdist arithmetic (single), False stays in double."""
    return _sdf_from_occupancy(ground_truth_map, cell_size, 3, single_precision, ctx)


def signedDistanceField2D(ground_truth_map, cell_size, single_precision=True, ctx=None):
    """matlab/+gpmp2/signedDistanceField2D.m on the device."""
    return _sdf_from_occupancy(ground_truth_map, cell_size, 2, single_precision, ctx)
