"""Secondary measurements: the other BASELINE.json configs (planar 2-link, planar 3-link + GP, Pose2MobileArm, WAM K=9)
through the host-buffer C-ABI call.  Prints one JSON line per config (not the headline bench)."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth


def run(name, model, sdf, st, pr, reps=3):
    a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
    G.batch_optimize(model, sdf, *a, st)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = G.batch_optimize(model, sdf, *a, st)
    dt = (time.perf_counter() - t0) / reps
    ks = G.default_context().last_kernel_stats()
    B = pr["init_traj"].shape[0]
    print(json.dumps({"config": name, "batch": B, "e2e_traj_per_s": B / dt, "kernel_ms": ks["kernel_ms"],
                      "kernel_traj_per_s": B / (ks["kernel_ms"] * 1e-3), "mean_iters": float(r["iters"].mean()),
                      "linearizations": ks["linearizations"], "solves": ks["solves"], "error_evals": ks["error_evals"]}))


run("config1 planar 2-link, OneObstacle 300x300, K=4, B=4096", synth.simple_two_links_arm(), synth.planar_dataset("OneObstacleDataset"),
    synth.bench_setting(2, total_time=10.0, cost_sigma=0.1, epsilon=0.1, inter=4), synth.planar_problems(4096, 2, seed=1))
run("config2 planar 3-link + GP, TwoObstacles, K=5, B=4096", synth.simple_three_links_arm(), synth.planar_dataset("TwoObstaclesDataset"),
    synth.bench_setting(3, total_time=10.0, cost_sigma=0.1, epsilon=0.2, inter=5), synth.planar_problems(4096, 3, seed=2))
run("config4 Pose2MobileArm 2-link, MobileMap1 500x500, K=5, B=16384", synth.mobile_two_links_arm(), synth.mobile_map(),
    synth.bench_setting(5, total_time=5.0, cost_sigma=0.1, epsilon=0.1), synth.mobile_problems(16384, seed=4, extent=3.5))
sdf = synth.wam_desk_dataset(300)
for B in (1024, 16384, 65536, 262144, 1048576):
    run("config5 WAM sweep K=5 B=%d" % B, synth.wam_arm(), sdf, synth.bench_setting(7), synth.wam_problems(B, seed=3))
run("WAM example setting K=9 (100 check points) B=65536", synth.wam_arm(), sdf, synth.bench_setting(7, inter=9), synth.wam_problems(65536, seed=3))
