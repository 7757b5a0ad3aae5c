"""Wrapper-interface functions that ride on existing device entry points: RobotModel::sphereCentersMat
(gpmp2/kinematics/RobotModel-inl.h:71-82) through gpmp2b_obstacle_errors, and the bare-Pose2 trajectory utilities
initPose2TrajStraightLine / interpolatePose2Traj (gpmp2/planner/TrajUtils.cpp:76-93, 239-275) as dof-3 Pose2Vector calls of
gpmp2b_init_straight_line / gpmp2b_interpolate_traj.  The device behaviour at these sizes was probed on a B200 first
(scripts/probe_facade_extras.py: <= 4e-15 against the oracle).  CPU tests: the Python glue with the three device entry
points replaced by the oracle; GPU tests: the real thing against the oracle."""
import numpy as np
import pytest

import gpmp2_b200 as G
from gpmp2_b200 import api, synth

S0, S1 = G.Pose2(0.3, -0.2, 0.4), G.Pose2(1.0, 2.0, -2.9)


def _oracle_backend(monkeypatch, oracle):
    """Stand-ins with the signatures of api.batch_obstacle_errors / batch_init_straight_line / batch_interpolate_traj."""
    def obstacle_errors(model, sdf, traj, setting, want_centers=True, ctx=None):
        return oracle.obstacle_errors(model, sdf, traj, setting, want_centers=want_centers)

    def init_line(start_conf, end_conf, total_step, lie=False, ctx=None):
        s = np.asarray(start_conf, dtype=np.float64)
        return oracle.init_straight_line(lie, s.shape[-1], total_step, s, end_conf)

    def interpolate(traj, dof, total_step, delta_t, inter_step, Qc=None, lie=False, start_index=0, end_index=None, ctx=None):
        return oracle.interpolate_traj(lie, dof, total_step, delta_t, Qc, inter_step, traj, start_index, end_index)

    monkeypatch.setattr(api, "batch_obstacle_errors", obstacle_errors)
    monkeypatch.setattr(api, "batch_init_straight_line", init_line)
    monkeypatch.setattr(api, "batch_interpolate_traj", interpolate)


def _check_sphere_centers(oracle):
    wam, q = synth.wam_arm(), np.linspace(-1.0, 1.0, 7)
    m = wam.sphereCentersMat(q)
    assert m.shape == (3, wam.nr_body_spheres())
    assert np.abs(m.T - oracle.sphere_centers(wam, q, want_J=False)[0]).max() < 1e-12
    mob = synth.mobile_two_links_arm()
    pq = G.Pose2Vector(G.Pose2(0.5, -1.0, 0.7), [0.3, -0.4])
    mm = mob.sphereCentersMat(pq)
    assert mm.shape == (3, mob.nr_body_spheres())
    assert np.abs(mm.T - oracle.sphere_centers(mob, pq.flat(), want_J=False)[0]).max() < 1e-12


def _close_mod_2pi(got, exp, n, tol=1e-12):
    """[x | v] wire vectors of n Pose2 states: equal up to a multiple of 2 pi in the headings."""
    d = (got - exp).reshape(2, n, 3)
    d[0, :, 2] = np.angle(np.exp(1j * d[0, :, 2]))
    return np.abs(d).max() < tol


def _check_pose2_trajectories(oracle):
    flat = lambda p: np.array([p.x(), p.y(), p.theta()])
    vals = G.initPose2TrajStraightLine(S0, S1, 4)
    assert vals.size() == 10 and isinstance(vals.at(G.symbol("x", 2)), G.Pose2)
    exp = oracle.init_straight_line(True, 3, 4, flat(S0), flat(S1))[0]
    got = np.concatenate([flat(vals.at(G.symbol("x", i))) for i in range(5)] + [vals.at(G.symbol("v", i)) for i in range(5)])
    assert _close_mod_2pi(got, exp, 5)
    # the reference's average velocity: (end - start) / total_step, theta difference not wrapped (TrajUtils.cpp:81-82)
    assert np.allclose(vals.at(G.symbol("v", 0)), (flat(S1) - flat(S0)) / 4.0)
    dense = G.interpolatePose2Traj(vals, np.eye(3), 0.5, 2, 1, 3)
    nout = (3 - 1) * 3 + 1
    assert dense.size() == 2 * nout and isinstance(dense.at(G.symbol("x", 0)), G.Pose2)
    expd = oracle.interpolate_traj(True, 3, 4, 0.5, np.eye(3), 2, exp, 1, 3)[0]
    gotd = np.concatenate([flat(dense.at(G.symbol("x", i))) for i in range(nout)] + [dense.at(G.symbol("v", i)) for i in range(nout)])
    assert _close_mod_2pi(gotd, expd, nout)
    assert np.allclose(flat(dense.at(G.symbol("x", 0))), flat(vals.at(G.symbol("x", 1))))     # starts at support state 1


def test_sphere_centers_mat_glue(monkeypatch, oracle):
    _oracle_backend(monkeypatch, oracle)
    _check_sphere_centers(oracle)


def test_pose2_trajectory_utilities_glue(monkeypatch, oracle):
    _oracle_backend(monkeypatch, oracle)
    _check_pose2_trajectories(oracle)


@pytest.mark.gpu
def test_sphere_centers_mat_device(oracle):
    _check_sphere_centers(oracle)


@pytest.mark.gpu
def test_pose2_trajectory_utilities_device(oracle):
    _check_pose2_trajectories(oracle)
