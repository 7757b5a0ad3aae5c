import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def limit_problem(conf, lie):
    """Two states at `conf` with zero velocity, priors at the states, an empty field: only the limit hinges contribute."""
    import numpy as np
    import gpmp2_b200 as G
    from gpmp2_b200 import synth
    model = synth.mobile_two_links_arm() if lie else synth.simple_two_links_arm()
    sdf = G.PlanarSDF([-20.0, -20.0], 1.0, np.full((40, 40), 1000.0))
    D = 5 if lie else 2
    st = G.TrajOptimizerSetting(D)
    st.set_total_step(1); st.set_total_time(1.0); st.set_obs_check_inter(0); st.setLM()
    st.set_flag_pos_limit(True)
    pad = [0.0, 0.0, 0.0] if lie else []
    st.set_joint_pos_limits_down(pad + [-5.0, -10.0]); st.set_joint_pos_limits_up(pad + [5.0, 10.0])
    st.set_pos_limit_thresh(pad + [2.0, 2.0]); st.set_pos_limit_model(np.ones(D))
    x = np.array(pad + list(conf), dtype=float)
    traj = np.concatenate([x, x, np.zeros(2 * D)])
    return model, sdf, st, x, traj


def limit_optimization_problem(conf):
    """testJointLimitFactorVector.cpp:71-157 as a two-state planner graph: weak priors (sigma 1000) at `conf` on both
    states, limit hinges with sigma 0.001, empty field, Gauss-Newton.  The states are symmetric, so the GP prior stays
    at zero error and each state solves the reference's one-variable problem."""
    import numpy as np
    model, sdf, st, x, traj = limit_problem(conf, False)
    st.setGaussNewton()
    st.set_conf_prior_model(1000.0); st.set_vel_prior_model(1000.0)
    st.set_pos_limit_model(0.001 * np.ones(2))
    st.set_max_iter(10); st.set_rel_thresh(0.0)
    return model, sdf, st, x, traj


def goal_ik_problem(g):
    """testGoalFactorArm.cpp:77-107 as a two-state planner graph: GoalFactorArm (sigma 0.1) on x_1 instead of the
    end-configuration prior, everything else made negligible (priors with sigma 1000, Qc = 1e6 I, empty field), LM as
    gpmp2::optimize runs it, straight-line initial values at q = 0."""
    import numpy as np
    import gpmp2_b200 as G
    o = g["optimization"]
    model = G.ArmModel(G.Arm(2, g["a"], g["alpha"], g["d"]), [G.BodySphere(1, 0.01, [0, 0, 0])])
    sdf = G.PlanarSDF([-20.0, -20.0], 1.0, np.full((40, 40), 1000.0))
    st = G.TrajOptimizerSetting(2)
    st.set_total_step(1); st.set_total_time(1.0); st.set_obs_check_inter(0); st.setLM()
    st.set_conf_prior_model(1000.0); st.set_vel_prior_model(1000.0); st.set_Qc_model(1e6 * np.eye(2))
    st.set_max_iter(100); st.set_rel_thresh(0.0)
    st.set_workspace_goal(o["goal"], o["cost_sigma"])
    start = np.asarray(o["qinit"], dtype=float)
    init = np.concatenate([start, start, np.zeros(4)])
    return model, sdf, st, start, start.copy(), init


def pose_ik_problem(o):
    """testGaussianPriorWorkspacePose.cpp:50-78 as a two-state planner graph: GaussianPriorWorkspacePoseArm (sigma 0.1) on
    x_1 instead of the end-configuration prior, everything else negligible (see goal_ik_problem)."""
    import numpy as np
    import gpmp2_b200 as G
    model = G.ArmModel(G.Arm(2, o["a"], o["alpha"], o["d"]), [G.BodySphere(1, 0.01, [0, 0, 0])])
    sdf = G.PlanarSDF([-20.0, -20.0], 1.0, np.full((40, 40), 1000.0))
    st = G.TrajOptimizerSetting(2)
    st.set_total_step(1); st.set_total_time(1.0); st.set_obs_check_inter(0); st.setLM()
    st.set_conf_prior_model(1000.0); st.set_vel_prior_model(1000.0); st.set_Qc_model(1e6 * np.eye(2))
    st.set_max_iter(100); st.set_rel_thresh(0.0)
    st.set_workspace_pose_goal(np.eye(3), o["des_t"], o["cost_sigma"])
    start = np.asarray(o["qinit"], dtype=float)
    init = np.concatenate([start, start, np.zeros(4)])
    return model, sdf, st, start, start.copy(), init
