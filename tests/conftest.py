import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The library sends LM problems of arms through the phase-kernel pipeline with the tensor-core solve only when the
    # batch is large (>= 8192 problems, dof >= 4: below that the one-kernel optimizer is faster).  The parity tests run
    # small batches, so the thresholds are lowered here: every LM test of an arm then exercises the pipeline.  The
    # one-kernel LM path keeps its coverage through tests/test_gpu_parity.py::test_fused_lm_kernel_matches_pipeline.
    os.environ.setdefault("GPMP2B_PK_MIN_BATCH", "1")
    os.environ.setdefault("GPMP2B_PK_MIN_DOF", "1")


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


# ---------------------------------------------------------------------------------------------
# optimizer-level parity with classification of every non-matching problem
# ---------------------------------------------------------------------------------------------
_PARITY_LOG = []


def _parity_log_path():
    return os.environ.get("GPMP2B_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "parity_fractions.json"))


def _current_test():
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]


def check_optimize_parity(oracle, G, model, sdf, st, pr, min_match, tol, label=None):
    import numpy as np
    from oracle.parity import oracle_sensitivity
    a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
    got = G.batch_optimize(model, sdf, *a, st)
    ref = oracle.batch_optimize(model, sdf, *a, st, nthreads=8)
    d = np.abs(got["traj"] - ref["traj"]).max(axis=1)
    same = (got["iters"] == ref["iters"]) & (d < tol)
    bad = np.nonzero(~same)[0]
    sens = oracle_sensitivity(model, sdf, st, pr, bad, ref, tol)
    unexplained = bad[~sens]
    frac = float(same.mean())
    rec = {"test": _current_test(), "label": label, "n": int(len(d)), "match_frac": frac, "min_match": min_match,
           "n_mismatch": int(len(bad)), "n_mismatch_oracle_sensitive": int(sens.sum()),
           "n_mismatch_unexplained": int(len(unexplained)),
           "worst_dev_matching_rad": float(d[same].max()) if same.any() else None,
           "worst_dev_all_rad": float(d.max()), "opt_type": int(st.opt_type) if hasattr(st, "opt_type") else None}
    _PARITY_LOG.append(rec)
    # every mismatch must be a rounding-decided branch of the algorithm itself, never a numeric difference
    assert len(unexplained) == 0, "problems %s differ from the oracle (worst %.3e rad) although the oracle is stable on them" % (
        unexplained.tolist(), d[unexplained].max())
    assert frac >= min_match, "only %.3f of problems within %g rad; worst %.3e" % (frac, tol, d.max())
    ok = same
    assert np.abs(got["error"][ok] / ref["error"][ok] - 1).max() < 1e-6   # trajectories agree to 1e-6 rad
    # bit 64 (ERR_INCREASED) is decided by `error > currentError`: at a converged Gauss-Newton step the two
    # errors agree to ~1e-15 relative and rounding picks the branch (either returned iterate is within 1e-14)
    assert ((got["status"][ok] & ~64) == (ref["status"][ok] & ~64)).all()
    assert np.allclose(got["coll_cost"][ok], ref["coll_cost"][ok], atol=1e-9)
    return frac, float(d.max())


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_LOG:
        return
    import json
    path = _parity_log_path()
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"tolerance_rad": 1e-6, "perturbation_rad": 1e-12, "records": _PARITY_LOG}, f, indent=1)
    except OSError:
        pass
