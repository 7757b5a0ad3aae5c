import sys
sys.path.insert(0, '.')
import numpy as np, gpmp2_b200 as G
from gpmp2_b200 import synth
wam = synth.wam_arm(); desk = synth.wam_desk_dataset(60); st = synth.bench_setting(7)
pr = synth.wam_problems(96, seed=3)
a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
r = G.batch_optimize(wam, desk, *a, st)
print("wam ok", r["iters"].mean(), np.isfinite(r["traj"]).all())
st2 = synth.bench_setting(7); st2.setDogleg(); r = G.batch_optimize(wam, desk, *a, st2); print("dogleg ok", r["iters"].mean())
m = synth.mobile_two_links_arm(); sdf = synth.mobile_map(); stm = synth.bench_setting(5, total_time=5.0, cost_sigma=0.1, epsilon=0.1)
pm = synth.mobile_problems(40, seed=4, extent=3.5)
r = G.batch_optimize(m, sdf, pm["start_conf"], pm["start_vel"], pm["end_conf"], pm["end_vel"], pm["init_traj"], stm)
print("mobile ok", r["iters"].mean())
H = G.batch_linearize(wam, desk, *a, st); print("lin ok", np.isfinite(H["Hdiag"]).all())
occ = np.zeros((20, 24, 28)); occ[5:9, 6:12, 10:15] = 1
print("edt ok", np.isfinite(G.signedDistanceField3D(occ, 0.05)).all())
