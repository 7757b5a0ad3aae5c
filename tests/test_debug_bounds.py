"""The in-tree memory-safety net (compute-sanitizer is not available on the GPU pool): libgpmp2b_dbg.so is the library
compiled with -DGPMP2B_DEBUG_BOUNDS (gpmp2_b200/csrc/kparams.h) -- every SDF cell index, every shared-memory store of
the tensor-core solve kernel and its assembly, and the work-list / scratch indices of the phase kernels are checked, an
index outside its array traps (= a CUDA error of the call).  Here it is driven at odd sizes over robot kinds, dofs,
optimizers and optional factors; the results must equal the release library's bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "gpmp2_b200", "csrc", "libgpmp2b_dbg.so")

DRIVER = r'''
import sys, json, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
import gpmp2_b200 as G
from gpmp2_b200 import synth
out = {}
def run(tag, model, sdf, st, pr):
    r = G.batch_optimize(model, sdf, pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"], st)
    assert np.isfinite(r["traj"]).all(), tag
    out[tag] = hashlib.sha256(np.ascontiguousarray(r["traj"]).tobytes() + np.ascontiguousarray(r["iters"]).tobytes()).hexdigest()
wam, desk = synth.wam_arm(), synth.wam_desk_dataset(40)
for steps in (1, 3, 10, 37):
    for K in (0, 5, 7, 24):
        if steps == 37 and K == 24:
            continue                                   # (1000 collision-checked configurations per trajectory: slow, nothing new)
        for opt in ("lm", "gn", "dogleg"):
            if opt != "lm" and (steps, K) not in ((3, 7), (10, 5)):
                continue
            st = synth.bench_setting(7, total_step=steps, inter=K, max_iter=4)
            if opt == "gn": st.setGaussNewton()
            if opt == "dogleg": st.setDogleg()
            run("wam_%%d_%%d_%%s" %% (steps, K, opt), wam, desk, st, synth.wam_problems(33, total_step=steps, seed=steps * 100 + K))
# limits + goal + self-collision variants (the EXTRA kernels), a field the robot partly leaves (out-of-range lookups)
st = synth.bench_setting(7, total_step=5, inter=3, max_iter=4)
st.set_flag_pos_limit(True); st.set_joint_pos_limits_up(1.5 * np.ones(7)); st.set_joint_pos_limits_down(-1.5 * np.ones(7))
st.set_pos_limit_thresh(0.01 * np.ones(7)); st.set_pos_limit_model(0.02 * np.ones(7))
run("wam_limits", wam, desk, st, synth.wam_problems(17, total_step=5, seed=5))
st = synth.bench_setting(7, total_step=5, inter=3, max_iter=4)
st.set_workspace_goal([0.4, 0.1, 0.5], 0.05); st.set_self_collision([[0, 9, 0.02, 0.1], [3, 12, 0.02, 0.1]])
run("wam_goal_self", wam, desk, st, synth.wam_problems(17, total_step=5, seed=6))
small = G.SignedDistanceField([-0.2, -0.2, -0.2], 0.05, np.full((9, 9, 9), 0.05))      # the arm reaches far outside it
run("wam_tiny_field", wam, small, synth.bench_setting(7, total_step=4, inter=2, max_iter=3), synth.wam_problems(9, total_step=4, seed=7))
# planar arms (2 and 3 dof, 2-D field) and the planar mobile manipulator
for name, B in (("planar2", 65), ("planar3gp", 65), ("mobile", 33)):
    for steps, K in ((1, 0), (3, 7), (10, 4)):
        cfg = synth.baseline_config(name, inter=K)
        st = cfg["setting"]; st.set_total_step(steps); st.set_max_iter(4)
        if name == "mobile":
            pr = synth.mobile_problems(B, total_step=steps, seed=cfg["seed"] + steps, extent=3.5)
        else:
            pr = synth.planar_problems(B, cfg["D"], total_step=steps, seed=cfg["seed"] + steps)
        run("%%s_%%d_%%d" %% (name, steps, K), cfg["model"], cfg["sdf"], st, pr)
# the other Pose2Vector robots (torso link, second arm, reversed actuator) through both LM paths and Gauss-Newton
for kind in ("two_arms", "vetlin_reversed", "vetlin_two_arms"):
    model = synth.other_mobile_robot(kind)
    for steps, K, opt in ((3, 7, "lm"), (10, 5, "lm"), (4, 2, "gn")):
        st = synth.bench_setting(model.dof(), total_time=5.0, total_step=steps, cost_sigma=0.1, epsilon=0.15, inter=K, max_iter=4)
        if opt == "gn": st.setGaussNewton()
        for fname, field in (("map", synth.mobile_map()), ("tiny", small)):
            run("%%s_%%d_%%d_%%s_%%s" %% (kind, steps, K, opt, fname), model, field, st,
                synth.other_mobile_problems(model, 21, seed=steps + K, total_step=steps, extent=3.5 if fname == "map" else 0.5))
print("RESULT " + json.dumps(out))
'''


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["GPMP2B_LIB"] = lib
    else:
        env.pop("GPMP2B_LIB", None)
    p = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert "GPMP2B_DEBUG_BOUNDS:" not in p.stdout + p.stderr, p.stdout[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_debug_library_is_built():
    assert os.path.exists(DBG), "run `make -C gpmp2_b200/csrc debug` (done by __graft_entry__.build())"


@pytest.mark.gpu
def test_debug_bounds_build_runs_clean_and_matches_release():
    dbg = _run(DBG)
    rel = _run(None)
    assert len(dbg) >= 48, sorted(dbg)
    assert dbg.keys() == rel.keys()
    bad = [k for k in dbg if dbg[k] != rel[k]]
    assert not bad, bad
