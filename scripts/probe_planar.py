"""One batch of BASELINE config 2 (planar 3-link arm + GP factors, B = 4096 x 8) for profiling."""
import sys
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
model = synth.simple_three_links_arm(); sdf = synth.planar_dataset("TwoObstaclesDataset")
st = synth.bench_setting(3, total_time=10.0, cost_sigma=0.1, epsilon=0.2, inter=5)
pr = synth.planar_problems(32768, 3, seed=2)
a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
for _ in range(2):
    r = G.batch_optimize(model, sdf, *a, st)
print(G.default_context().last_kernel_stats(), r["iters"].mean(), flush=True)
