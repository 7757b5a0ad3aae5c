"""How much of the kernel time is per-warp latency vs. contention: time one wave at 1..7 warps per SM."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
sdf = synth.wam_desk_dataset(300); model = synth.wam_arm(); st = synth.bench_setting(7)
ctx = G.default_context()
for B in (37, 148, 296, 592, 1036, 2072):
    pr = synth.wam_problems(B, seed=3)
    a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
    for _ in range(3):
        G.batch_optimize(model, sdf, *a, st)
    ks = ctx.last_kernel_stats()
    print(json.dumps({"B": B, "warps_per_sm": B / 148.0, "kernel_ms": ks["kernel_ms"], "lin": ks["linearizations"],
                      "kcycles_per_lm_iter_per_warp": ks["kernel_ms"] * 1e-3 * 1.965e9 / (ks["linearizations"] / B) / 1e3 / max(1, B / 1036)}))
