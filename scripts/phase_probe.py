"""Per-phase clock64 totals of a GPMP2B_PHASE_TIMING build (scripts/build_variant_pt.sh); select it with GPMP2B_LIB."""
import sys
sys.path.insert(0, '.')
import gpmp2_b200 as G
from gpmp2_b200 import synth
sdf = synth.wam_desk_dataset(300); model = synth.wam_arm(); st = synth.bench_setting(7)
ctx = G.default_context()
for B in (1036, 65536):
    pr = synth.wam_problems(B, seed=3)
    a = (pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"])
    for _ in range(2):
        G.batch_optimize(model, sdf, *a, st)
    print("B", B, ctx.last_kernel_stats(), flush=True)
