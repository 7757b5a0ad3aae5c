"""Secondary measurement: the optional factors of hand-built graphs (workspace goal, self-collision) on the
headline WAM workload, through the host-buffer C-ABI call.  One JSON line per variant (not the headline bench)."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth

PAIRS = [[0, 12, 0.45, 0.05], [1, 15, 0.50, 0.1], [3, 14, 0.30, 0.02], [5, 13, 0.25, 0.05], [2, 9, 0.2, 0.05]]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wam, sdf = synth.wam_arm(), synth.wam_desk_dataset(300)
pr = synth.wam_problems(B, seed=3)
a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
for name, goal, pairs in (("plain", False, None), ("goal", True, None), ("self-collision x5", False, PAIRS),
                          ("goal + self-collision x5", True, PAIRS), ("self-collision x32", False, (PAIRS * 7)[:32])):
    st = synth.bench_setting(7)
    if goal:
        st.set_workspace_goal([0.6, 0.1, 0.3], 0.02)
    st.set_self_collision(pairs)
    G.batch_optimize(wam, sdf, *a, st)
    t0 = time.perf_counter()
    for _ in range(3):
        r = G.batch_optimize(wam, sdf, *a, st)
    dt = (time.perf_counter() - t0) / 3
    ks = G.default_context().last_kernel_stats()
    print(json.dumps({"config": "WAM K=5 B=%d, %s" % (B, name), "e2e_traj_per_s": B / dt, "kernel_ms": ks["kernel_ms"],
                      "kernel_traj_per_s": B / (ks["kernel_ms"] * 1e-3), "mean_iters": float(r["iters"].mean()),
                      "linearizations": ks["linearizations"], "solves": ks["solves"], "error_evals": ks["error_evals"]}))
