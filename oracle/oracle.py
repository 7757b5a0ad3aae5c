"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  It reuses the struct layouts of the product header
(gpmp2_b200/_abi.py mirrors include/gpmp2b.h) but never the product's compute.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gpmp2_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(
            os.path.join(_HERE, "gpmp2_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def forward_kinematics(model, conf, want_J=True):
    conf = _f64(conf)
    L = model.fk_model().nr_links() if model.kind != _abi.ROBOT_ARM else model.dof()
    D = model.dof()
    poses = np.zeros((L, 4, 4))
    J = np.zeros((L, 6, D)) if want_J else None
    rc = lib().orc_forward_kinematics(C.byref(model.desc), _p(conf), _p(poses), _p(J))
    assert rc == 0
    return poses, J


def sphere_centers(model, conf, want_J=True):
    conf = _f64(conf)
    S, D = model.nr_body_spheres(), model.dof()
    c = np.zeros((S, 3))
    J = np.zeros((S, 3, D)) if want_J else None
    assert lib().orc_sphere_centers(C.byref(model.desc), _p(conf), _p(c), _p(J)) == 0
    return c, J


def sdf_query(sdf, p):
    """-> (in_range, dist, grad)"""
    p = _f64(p)
    d = C.c_double()
    g = np.zeros(sdf.ndim)
    rc = lib().orc_sdf_query(C.byref(sdf.desc), _p(p), C.byref(d), _p(g))
    assert rc >= 0
    return bool(rc), d.value, g


def obstacle_factor(model, sdf, conf, epsilon, want_H=True):
    conf = _f64(conf)
    S, D = model.nr_body_spheres(), model.dof()
    e = np.zeros(S)
    H = np.zeros((S, D)) if want_H else None
    assert lib().orc_obstacle_factor(C.byref(model.desc), C.byref(sdf.desc), _p(conf), C.c_double(epsilon),
                                     _p(e), _p(H)) == 0
    return e, H


def goal_factor(model, conf, goal, link=-1, want_H=True):
    """GoalFactorArm / GaussianPriorWorkspacePosition: (e[3], H[3][dof]); link -1 = last joint frame."""
    D = model.dof()
    e = np.zeros(3)
    H = np.zeros((3, D)) if want_H else None
    assert lib().orc_goal_factor(C.byref(model.desc), _p(_f64(conf)), C.c_int(int(link)), _p(_f64(goal)), _p(e), _p(H)) == 0
    return e, H


def orientation_factor(model, conf, des_R, link=-1, want_H=True):
    """GaussianPriorWorkspaceOrientation: (e[3], H[3][dof]); des_R 3x3 desired rotation; link -1 = last link frame."""
    D = model.dof()
    e = np.zeros(3)
    H = np.zeros((3, D)) if want_H else None
    R = _f64(np.asarray(des_R, dtype=np.float64).reshape(3, 3))
    assert lib().orc_orientation_factor(C.byref(model.desc), _p(_f64(conf)), C.c_int(int(link)), _p(R), _p(e), _p(H)) == 0
    return e, H


def pose_factor(model, conf, des_R, des_t, link=-1, want_H=True):
    """GaussianPriorWorkspacePose: (e[6] = [omega; u], H[6][dof]); link -1 = last link frame."""
    D = model.dof()
    e = np.zeros(6)
    H = np.zeros((6, D)) if want_H else None
    R = _f64(np.asarray(des_R, dtype=np.float64).reshape(3, 3))
    assert lib().orc_pose_factor(C.byref(model.desc), _p(_f64(conf)), C.c_int(int(link)), _p(R), _p(_f64(des_t)), _p(e), _p(H)) == 0
    return e, H


def self_collision_factor(model, conf, data, want_H=True):
    """SelfCollisionArm::evaluateError: (e[n], H[n][dof]) for data rows (sphere A, sphere B, epsilon, sigma)."""
    data = _f64(np.asarray(data, dtype=np.float64).reshape(-1, 4))
    n, D = data.shape[0], model.dof()
    e = np.zeros(n)
    H = np.zeros((n, D)) if want_H else None
    assert lib().orc_self_collision_factor(C.byref(model.desc), _p(_f64(conf)), C.c_int(n), _p(data), _p(e), _p(H)) == 0
    return e, H


def obstacle_gp_factor(model, sdf, Qc, delta_t, tau, x1, v1, x2, v2, epsilon, want_H=True):
    S, D = model.nr_body_spheres(), model.dof()
    e = np.zeros(S)
    H = np.zeros((4, S, D)) if want_H else None
    Qc = None if Qc is None else _f64(Qc)
    assert lib().orc_obstacle_gp_factor(C.byref(model.desc), C.byref(sdf.desc), _p(Qc), C.c_double(delta_t),
                                        C.c_double(tau), _p(_f64(x1)), _p(_f64(v1)), _p(_f64(x2)), _p(_f64(v2)),
                                        C.c_double(epsilon), _p(e), _p(H)) == 0
    return e, H


def gp_interpolate(dof, lie, Qc, delta_t, tau, x1, v1, x2, v2, want_H=True):
    out = np.zeros(dof)
    H = np.zeros((4, dof, dof)) if want_H else None
    Qc = None if Qc is None else _f64(Qc)
    assert lib().orc_gp_interpolate(dof, int(lie), _p(Qc), C.c_double(delta_t), C.c_double(tau), _p(_f64(x1)),
                                    _p(_f64(v1)), _p(_f64(x2)), _p(_f64(v2)), _p(out), _p(H)) == 0
    return out, H


def gp_prior(dof, lie, delta_t, x1, v1, x2, v2, want_H=True):
    e = np.zeros(2 * dof)
    H = np.zeros((4, 2 * dof, dof)) if want_H else None
    assert lib().orc_gp_prior(dof, int(lie), C.c_double(delta_t), _p(_f64(x1)), _p(_f64(v1)), _p(_f64(x2)),
                              _p(_f64(v2)), _p(e), _p(H)) == 0
    return e, H


def gp_lambda_psi(dof, Qc, delta_t, tau):
    L, P = np.zeros((2 * dof, 2 * dof)), np.zeros((2 * dof, 2 * dof))
    Qc = None if Qc is None else _f64(Qc)
    assert lib().orc_gp_lambda_psi(dof, _p(Qc), C.c_double(delta_t), C.c_double(tau), _p(L), _p(P)) == 0
    return L, P


def _batch_args(setting, start_conf, start_vel, end_conf, end_vel, traj):
    D, N = setting.dof, setting.total_step + 1
    traj = _f64(traj).reshape(-1, 2 * N * D)
    B = traj.shape[0]
    sc, sv = _f64(start_conf).reshape(B, D), _f64(start_vel).reshape(B, D)
    ec, ev = _f64(end_conf).reshape(B, D), _f64(end_vel).reshape(B, D)
    return B, D, N, sc, sv, ec, ev, traj


def linearize(model, sdf, start_conf, start_vel, end_conf, end_vel, traj, setting, want_dense=False):
    B, D, N, sc, sv, ec, ev, traj = _batch_args(setting, start_conf, start_vel, end_conf, end_vel, traj)
    b = 2 * D
    Hd, Ho = np.zeros((B, N, b, b)), np.zeros((B, N - 1, b, b))
    g, err = np.zeros((B, N, b)), np.zeros(B)
    dense = np.zeros((B, N * b, N * b)) if want_dense else None
    s, keep = setting.pack()
    rc = lib().orc_linearize(C.byref(model.desc), C.byref(sdf.desc), C.byref(s), C.c_int64(B), _p(sc), _p(sv),
                             _p(ec), _p(ev), _p(traj), _p(Hd), _p(Ho), _p(g), _p(err), _p(dense))
    assert rc == 0
    return {"Hdiag": Hd, "Hoff": Ho, "g": g, "error": err, "dense_H": dense}


def obstacle_errors(model, sdf, traj, setting, want_centers=True):
    D, N, K = setting.dof, setting.total_step + 1, setting.obs_check_inter
    traj = _f64(traj).reshape(-1, 2 * N * D)
    B = traj.shape[0]
    S = model.nr_body_spheres()
    Cn = N + (N - 1) * K
    err = np.zeros((B, Cn, S))
    ctr = np.zeros((B, Cn, S, 3)) if want_centers else None
    s, keep = setting.pack()
    assert lib().orc_obstacle_errors(C.byref(model.desc), C.byref(sdf.desc), C.byref(s), C.c_int64(B), _p(traj),
                                     _p(err), _p(ctr)) == 0
    return {"err": err, "centers": ctr}


def batch_optimize(model, sdf, start_conf, start_vel, end_conf, end_vel, init_traj, setting, nthreads=1,
                   dense=False):
    B, D, N, sc, sv, ec, ev, traj = _batch_args(setting, start_conf, start_vel, end_conf, end_vel, init_traj)
    out = np.zeros_like(traj)
    err, cc = np.zeros(B), np.zeros(B)
    iters, status = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    counts = np.zeros((B, 3), dtype=np.int64)
    s, keep = setting.pack()
    rc = lib().orc_batch_optimize(C.byref(model.desc), C.byref(sdf.desc), C.byref(s), C.c_int64(B), _p(sc), _p(sv),
                                  _p(ec), _p(ev), _p(traj), _p(out), _p(err), _p(cc), _p(iters), _p(status),
                                  _p(counts), int(nthreads), int(dense))
    assert rc == 0, "oracle batch_optimize failed"
    return {"traj": out, "error": err, "coll_cost": cc, "iters": iters, "status": status, "counts": counts}


def collision_cost(model, sdf, traj, setting):
    D, N = setting.dof, setting.total_step + 1
    traj = _f64(traj).reshape(-1, 2 * N * D)
    out = np.zeros(traj.shape[0])
    s, keep = setting.pack()
    assert lib().orc_collision_cost(C.byref(model.desc), C.byref(sdf.desc), C.byref(s), C.c_int64(traj.shape[0]),
                                    _p(traj), _p(out)) == 0
    return out


def graph_error(model, sdf, start_conf, start_vel, end_conf, end_vel, traj, setting):
    B, D, N, sc, sv, ec, ev, traj = _batch_args(setting, start_conf, start_vel, end_conf, end_vel, traj)
    out = np.zeros(B)
    s, keep = setting.pack()
    assert lib().orc_graph_error(C.byref(model.desc), C.byref(sdf.desc), C.byref(s), C.c_int64(B), _p(sc), _p(sv),
                                 _p(ec), _p(ev), _p(traj), _p(out)) == 0
    return out


def init_straight_line(lie, dof, total_step, start, end):
    """initArmTrajStraightLine / initPose2VectorTrajStraightLine for B (start, end) pairs -> [B][2*N*dof]."""
    start, end = _f64(start).reshape(-1, dof), _f64(end).reshape(-1, dof)
    B = start.shape[0]
    out = np.zeros((B, 2 * (total_step + 1) * dof))
    assert lib().orc_init_straight_line(int(lie), dof, total_step, C.c_int64(B), _p(start), _p(end), _p(out)) == 0
    return out


def interpolate_traj(lie, dof, total_step, delta_t, Qc, inter_step, traj, start_index=0, end_index=None):
    """interpolateArmTraj / interpolatePose2MobileArmTraj -> [B][2*Nout*dof]."""
    end_index = total_step if end_index is None else end_index
    traj = _f64(traj).reshape(-1, 2 * (total_step + 1) * dof)
    B = traj.shape[0]
    nout = (end_index - start_index) * (inter_step + 1) + 1
    out = np.zeros((B, 2 * nout * dof))
    Qc = None if Qc is None else _f64(Qc)
    assert lib().orc_interpolate_traj(int(lie), dof, total_step, C.c_double(delta_t), _p(Qc), inter_step, start_index,
                                      end_index, C.c_int64(B), _p(traj), _p(out)) == 0
    return out


def pose2_op(name, a, b=None):
    """Pose2 compose / between / inverse of the oracle (poses as (x, y, theta))."""
    out = np.zeros(3)
    fn = getattr(lib(), "orc_pose2_" + name)
    if b is None:
        assert fn(_p(_f64(a)), _p(out)) == 0
    else:
        assert fn(_p(_f64(a)), _p(_f64(b)), _p(out)) == 0
    return out
