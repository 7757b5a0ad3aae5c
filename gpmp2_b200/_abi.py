"""ctypes mirror of include/gpmp2b.h (struct layouts + function prototypes) and the library loader.

The product path has NO CPU fallback: `load_library()` raises if the CUDA library is missing, and
every compute entry point needs a CUDA device (gpmp2b_create fails loudly without one).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgpmp2b.so")

MAX_DOF = 8
MAX_SPHERES = 64

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4

ST_CONVERGED_ABS, ST_CONVERGED_REL, ST_MAX_ITER = 1, 2, 4
ST_LAMBDA_MAXED, ST_SOLVE_FAILED, ST_ERROR_TOL, ST_ERR_INCREASED = 8, 16, 32, 64

ROBOT_ARM, ROBOT_POSE2_MOBILE_ARM, ROBOT_POSE2_MOBILE_2ARMS, ROBOT_POSE2_MOBILE_VETLIN_ARM, ROBOT_POSE2_MOBILE_VETLIN_2ARMS = 0, 1, 2, 3, 4
OPT_GAUSS_NEWTON, OPT_LM, OPT_DOGLEG = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class RobotDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("arm_dof", C.c_int32), ("n_spheres", C.c_int32), ("reserved_", C.c_int32),
        ("a", c_double_p), ("alpha", c_double_p), ("d", c_double_p), ("theta_bias", c_double_p),
        ("base_pose", C.c_double * 16),
        ("sphere_link", c_int32_p), ("sphere_radius", c_double_p), ("sphere_center", c_double_p),
        ("arm2_dof", C.c_int32), ("reverse_linact", C.c_int32),
        ("base_pose2", C.c_double * 16), ("base_pose3", C.c_double * 16),
    ]


class SdfDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("nz", C.c_int32),
        ("origin", C.c_double * 3), ("cell_size", C.c_double), ("data", c_double_p),
    ]


class Setting(C.Structure):
    _fields_ = [
        ("dof", C.c_int32), ("total_step", C.c_int32), ("total_time", C.c_double),
        ("conf_prior_sigma", C.c_double), ("vel_prior_sigma", C.c_double),
        ("flag_pos_limit", C.c_int32), ("flag_vel_limit", C.c_int32),
        ("joint_pos_limits_up", c_double_p), ("joint_pos_limits_down", c_double_p),
        ("vel_limits", c_double_p), ("pos_limit_thresh", c_double_p), ("vel_limit_thresh", c_double_p),
        ("pos_limit_sigma", c_double_p), ("vel_limit_sigma", c_double_p),
        ("epsilon", C.c_double), ("cost_sigma", C.c_double),
        ("obs_check_inter", C.c_int32), ("opt_type", C.c_int32),
        ("Qc", c_double_p),
        ("opt_verbosity", C.c_int32), ("final_iter_no_increase", C.c_int32),
        ("rel_thresh", C.c_double), ("max_iter", C.c_int32), ("reserved_", C.c_int32),
        ("goal_enabled", C.c_int32), ("goal_link", C.c_int32), ("goal_keep_end_prior", C.c_int32), ("reserved2_", C.c_int32),
        ("goal_sigma", C.c_double), ("goal_pos", C.c_double * 3),
        ("n_self_collision", C.c_int32), ("reserved3_", C.c_int32), ("self_collision_data", c_double_p),
        ("vehicle_dynamics_sigma", C.c_double),
        ("orient_enabled", C.c_int32), ("orient_link", C.c_int32), ("orient_state_first", C.c_int32),
        ("orient_state_last", C.c_int32), ("orient_sigma", C.c_double), ("orient_R", C.c_double * 9),
        ("goal_R", C.c_double * 9),
        ("goal_pos_batch", C.c_void_p), ("goal_R_batch", C.c_void_p), ("orient_R_batch", C.c_void_p),
        ("fix_enabled", C.c_int32), ("fix_state_index", C.c_int32), ("fix_conf", C.c_void_p), ("fix_vel", C.c_void_p),
    ]


def dptr(arr):
    """double* of a C-contiguous float64 numpy array (or NULL for None)."""
    if arr is None:
        return None
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(c_double_p)


def iptr(arr):
    if arr is None:
        return None
    assert arr.dtype == np.int32 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(c_int32_p)


# every symbol include/gpmp2b.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "gpmp2b_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "gpmp2b_destroy": (None, [C.c_void_p]),
    "gpmp2b_last_error": (C.c_char_p, [C.c_void_p]),
    "gpmp2b_version": (C.c_char_p, []),
    "gpmp2b_robot_upload": (C.c_int, [C.c_void_p, C.POINTER(RobotDesc), C.POINTER(C.c_void_p)]),
    "gpmp2b_robot_free": (None, [C.c_void_p, C.c_void_p]),
    "gpmp2b_sdf_upload": (C.c_int, [C.c_void_p, C.POINTER(SdfDesc), C.POINTER(C.c_void_p)]),
    "gpmp2b_sdf_from_occupancy": (C.c_int, [C.c_void_p, C.POINTER(SdfDesc), C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "gpmp2b_sdf_free": (None, [C.c_void_p, C.c_void_p]),
    "gpmp2b_batch_optimize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Setting), C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_batch_optimize_multi": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                              C.POINTER(Setting), C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "gpmp2b_collision_cost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Setting), C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_linearize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Setting), C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_obstacle_errors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Setting), C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_init_straight_line": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_interpolate_traj": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_select_best": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "gpmp2b_measure_peaks": (C.c_int, [C.c_void_p, c_double_p]),
    "gpmp2b_launch_count": (C.c_int64, [C.c_void_p]),
    "gpmp2b_last_kernel_stats": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_int64_p, c_int64_p]),
}

_lib = None


def load_library(path=None):
    """dlopen the CUDA library and bind every prototype.  Fails loudly if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("GPMP2B_LIB") or LIB_PATH   # GPMP2B_LIB: developer override (kernel A/B experiments)
    if not os.path.exists(p):
        raise RuntimeError(
            "gpmp2_b200: CUDA library %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
