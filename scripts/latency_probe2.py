"""I-cache hypothesis: identical problems in every warp (all warps execute the same instruction sequence in near
lockstep) vs. distinct problems, at 7 warps per SM."""
import sys, json
import numpy as np
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
sdf = synth.wam_desk_dataset(300); model = synth.wam_arm(); st = synth.bench_setting(7)
ctx = G.default_context()
for B in (592, 1036):
    for same in (False, True):
        pr = synth.wam_problems(B, seed=3)
        if same:
            for k in pr: pr[k][:] = pr[k][0]
        a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
        for _ in range(3):
            r = G.batch_optimize(model, sdf, *a, st)
        ks = ctx.last_kernel_stats()
        print(json.dumps({"B": B, "identical_problems": same, "kernel_ms": ks["kernel_ms"], "lin": ks["linearizations"], "solves": ks["solves"],
                          "us_per_lm_iter_per_warp": ks["kernel_ms"] * 1e3 / (ks["linearizations"] / B)}))
