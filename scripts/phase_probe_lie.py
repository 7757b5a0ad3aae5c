"""Per-phase clock64 totals of the Pose2Vector kernel (scripts/build_variant_pt_lie.sh); select the build with GPMP2B_LIB."""
import sys
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
model = synth.mobile_two_links_arm(); sdf = synth.mobile_map()
st = synth.bench_setting(5, total_time=5.0, cost_sigma=0.1, epsilon=0.1)
pr = synth.mobile_problems(16384, seed=4, extent=3.5)
a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
for _ in range(2):
    r = G.batch_optimize(model, sdf, *a, st)
print(G.default_context().last_kernel_stats(), "mean iters", r["iters"].mean(), flush=True)
