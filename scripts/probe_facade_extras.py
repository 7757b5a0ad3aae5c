"""One-off GPU probe (no torch import): do the existing entry points serve (1) sphereCentersMat(conf) through
gpmp2b_obstacle_errors with total_step = 1, K = 0 and a dummy field, (2) bare Pose2 trajectories as dof-3 Pose2Vector states
in gpmp2b_init_straight_line / gpmp2b_interpolate_traj?  Compared with the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpmp2_b200 as G
from gpmp2_b200 import synth
from oracle import oracle as O

res = {}
model = synth.wam_arm()
q = np.linspace(-1.0, 1.0, 7)
ref_c, _ = O.sphere_centers(model, q, want_J=False)
s0 = np.array([0.3, -0.2, 0.4]); s1 = np.array([1.0, 2.0, -2.9])
to = O.init_straight_line(True, 3, 4, s0, s1)
do = O.interpolate_traj(True, 3, 4, 0.5, np.eye(3), 2, to)
mm = synth.mobile_two_links_arm()
pq = np.array([0.5, -1.0, 0.7, 0.3, -0.4])
ref_m, _ = O.sphere_centers(mm, pq, want_J=False)
print("oracle side ok", flush=True)

sdf = G.SignedDistanceField([0, 0, 0], 1.0, np.zeros((2, 2, 2)))
st = G.TrajOptimizerSetting(7); st.set_total_step(1); st.set_obs_check_inter(0)
c = G.batch_obstacle_errors(model, sdf, np.concatenate([q, q, np.zeros(14)]), st)["centers"][0, 0]
res["centers_wam"] = float(np.abs(c - ref_c).max())
stm = G.TrajOptimizerSetting(5); stm.set_total_step(1); stm.set_obs_check_inter(0)
cm = G.batch_obstacle_errors(mm, sdf, np.concatenate([pq, pq, np.zeros(10)]), stm)["centers"][0, 0]
res["centers_mobile"] = float(np.abs(cm - ref_m).max())
t = G.batch_init_straight_line(s0, s1, 4, lie=True)
res["init_pose2"] = float(np.abs(t - to).max())
d = G.batch_interpolate_traj(t, 3, 4, 0.5, 2, Qc=np.eye(3), lie=True)
res["interp_pose2"] = float(np.abs(d - do).max())
print(res, flush=True)
open("gpurun_out/probe_facade_extras.txt", "w").write(repr(res) + "\n")
