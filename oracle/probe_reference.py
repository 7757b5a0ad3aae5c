"""Probe for a REAL reference build (upstream gpmp2 + GTSAM): the only thing that can pin the optimizer level
(gpmp2::optimize over GTSAM's LevenbergMarquardtOptimizer, gpmp2/planner/BatchTrajOptimizer.cpp:212-308) bit for bit.

Test infrastructure (like everything under oracle/): run by __graft_entry__.build() and `bench.py --impl reference`.
Looks for, in this order,
  1. an importable `gpmp2` Python toolbox with `gtsam` (the reference's Cython wrap, gpmp2/gpmp2_python),
  2. a pip-installed copy under baseline/_ref,
and, if one is there, optimizes 64 WAM problems of the bench workload with BatchTrajOptimize3DArm and compares the
trajectories with the oracle's (tolerance 1e-6 rad).  The outcome -- found or not, and the differences -- is written
to oracle/_ref/probe.json (git-ignored) and returned.

In this image neither exists (no cmake-built GTSAM, no Boost / Eigen): the probe reports "absent" and the
optimizer level stays pinned only to the reference's small known answers."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find():
    tried = []
    for extra in (None, os.path.join(ROOT, "baseline", "_ref")):
        if extra and extra not in sys.path and os.path.isdir(extra):
            sys.path.insert(0, extra)
        try:
            gtsam = importlib.import_module("gtsam")
            gpmp2 = importlib.import_module("gpmp2")
            if hasattr(gpmp2, "BatchTrajOptimize3DArm"):
                return gtsam, gpmp2, tried
            tried.append("gpmp2 module without BatchTrajOptimize3DArm at %s" % getattr(gpmp2, "__file__", "?"))
        except Exception as e:   # noqa: BLE001 (ImportError, or a broken native module)
            tried.append("%s: %s" % (extra or "sys.path", e))
    return None, None, tried


def probe(n_problems=64, write=True):
    gtsam, gpmp2, tried = _find()
    out = {"found": gtsam is not None, "tried": tried}
    if gtsam is not None:
        out.update(_compare(gtsam, gpmp2, n_problems))
    if write:
        d = os.path.join(ROOT, "oracle", "_ref")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "probe.json"), "w") as f:
            json.dump(out, f, indent=1)
    return out


def _compare(gtsam, gpmp2, n):
    """The reference's own planner on the bench problems against the oracle (only reached with a real build)."""
    import numpy as np
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from gpmp2_b200 import synth
    from oracle import oracle as O
    cfg = synth.baseline_config("wam", sdf_cells=300)
    st, model, sdf = cfg["setting"], cfg["model"], cfg["sdf"]
    pr = cfg["problems"](n, cfg["seed"])
    D, N = st.dof, st.total_step + 1
    # reference objects through the wrap (names of gpmp2/gpmp2.h, the reference's interface file)
    arm = gpmp2.Arm(D, np.asarray(model.fk_model().a()), np.asarray(model.fk_model().alpha()), np.asarray(model.fk_model().d()),
                    gtsam.Pose3(), np.asarray(model.fk_model().theta_bias()))
    spheres = gpmp2.BodySphereVector()
    for s in model.spheres():
        spheres.push_back(gpmp2.BodySphere(s.link_id, s.radius, gtsam.Point3(*s.center)))
    ref_model = gpmp2.ArmModel(arm, spheres)
    wire = np.asarray(sdf._wire)
    ref_sdf = gpmp2.SignedDistanceField(gtsam.Point3(*sdf._origin), sdf._cell, wire.shape[2], wire.shape[1], wire.shape[0])
    for z in range(wire.shape[0]):
        ref_sdf.initFieldData(z, np.asfortranarray(wire[z].T))
    rs = gpmp2.TrajOptimizerSetting(D)
    rs.set_total_step(st.total_step); rs.set_total_time(st.total_time); rs.set_epsilon(st.epsilon); rs.set_cost_sigma(st.cost_sigma)
    rs.set_obs_check_inter(st.obs_check_inter); rs.set_conf_prior_model(st.conf_prior_sigma); rs.set_vel_prior_model(st.vel_prior_sigma)
    rs.set_Qc_model(np.asarray(st.Qc).reshape(D, D)); rs.setLM(); rs.set_max_iter(st.max_iter); rs.set_rel_thresh(st.rel_thresh)
    exp = O.batch_optimize(model, sdf, pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"], st)
    worst = 0.0
    for k in range(n):
        init = gtsam.Values()
        t = pr["init_traj"][k]
        for i in range(N):
            init.insert(gtsam.symbol(ord("x"), i), t[i * D:(i + 1) * D])
            init.insert(gtsam.symbol(ord("v"), i), t[(N + i) * D:(N + i + 1) * D])
        res = gpmp2.BatchTrajOptimize3DArm(ref_model, ref_sdf, pr["start_conf"][k], pr["start_vel"][k], pr["end_conf"][k],
                                           pr["end_vel"][k], init, rs)
        got = np.concatenate([res.atVector(gtsam.symbol(ord("x"), i)) for i in range(N)] +
                             [res.atVector(gtsam.symbol(ord("v"), i)) for i in range(N)])
        worst = max(worst, float(np.abs(got - exp["traj"][k]).max()))
    return {"n": n, "max_abs_rad_vs_oracle": worst, "within_1e-6": worst <= 1e-6}


if __name__ == "__main__":
    print(json.dumps(probe(), indent=1))
