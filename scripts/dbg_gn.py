import sys, numpy as np
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
from oracle import oracle as O
wam = synth.wam_arm(); desk = synth.wam_desk_dataset(100)
st2 = synth.bench_setting(7, max_iter=6); st2.setGaussNewton()
pr = synth.wam_problems(48, mode="random", seed=34)
a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
got = G.batch_optimize(wam, desk, *a, st2); ref = O.batch_optimize(wam, desk, *a, st2, nthreads=8)
d = np.abs(got["traj"] - ref["traj"]).max(axis=1)
for i in range(48):
    if got["status"][i] != ref["status"][i] or got["iters"][i] != ref["iters"][i] or not d[i] < 1e-6:
        print(i, got["status"][i], ref["status"][i], got["iters"][i], ref["iters"][i], d[i], got["error"][i], ref["error"][i])
print("done")
