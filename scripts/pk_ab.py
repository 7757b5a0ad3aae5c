"""A/B of the phase-kernel pipeline against the fused optimizer kernel: run with GPMP2B_PK=0 and GPMP2B_PK=1 (two
processes, argv[1] = output .npz), then `compare a.npz b.npz` checks bit-identity of every output."""
import sys
import numpy as np
sys.path.insert(0, '.')


def run(out):
    import gpmp2_b200 as G
    from gpmp2_b200 import synth
    res = {}
    cases = [("wam", 4096, {}), ("wam", 1, {}), ("wam", 37, {"inter": 9}), ("wam", 64, {"inter": 0}), ("planar2", 512, {}),
             ("planar3gp", 512, {}), ("planar3gp", 64, {"inter": 7})]
    import os
    if os.environ.get("PK_AB_WAM_ONLY"):    # variant libraries only carry rebuilt 7-DOF kernels
        cases = [c for c in cases if c[0] == "wam"]
    for name, B, kw in cases:
        cfg = synth.baseline_config(name, sdf_cells=100, **kw)
        pr = cfg["problems"](B, cfg["seed"])
        st = cfg["setting"]
        for variant in ("default", "relthresh", "limits"):
            if variant == "relthresh":
                st.set_rel_thresh(1e-2); st.set_max_iter(30)
            if variant == "limits":
                D = cfg["D"]
                st.set_flag_pos_limit(True); st.set_joint_pos_limits_up(2.0 * np.ones(D)); st.set_joint_pos_limits_down(-2.0 * np.ones(D))
                st.set_pos_limit_thresh(0.01 * np.ones(D)); st.set_pos_limit_model(0.02 * np.ones(D))
            r = G.batch_optimize(cfg["model"], cfg["sdf"], pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"],
                                 pr["init_traj"], st)
            ks = G.default_context().last_kernel_stats()
            key = "%s_%d_%s_%s" % (name, B, "_".join("%s%s" % kv for kv in kw.items()), variant)
            for k in ("traj", "error", "coll_cost", "iters", "status"):
                res[key + ":" + k] = r[k]
            res[key + ":counts"] = np.array([ks["linearizations"], ks["solves"], ks["error_evals"]])
            print(key, "kernel_ms %.3f" % ks["kernel_ms"], "iters %.2f" % r["iters"].mean(), ks, flush=True)
    np.savez(out, **res)


def compare(a, b):
    A, Bz = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        same = np.array_equal(A[k], Bz[k], equal_nan=True) if A[k].dtype.kind == "f" else np.array_equal(A[k], Bz[k])
        if not same:
            bad += 1
            d = np.abs(A[k].astype(float) - Bz[k].astype(float))
            big = int((d > 1e-6).sum()) if A[k].dtype.kind == "f" else int((d > 0).sum())
            print("DIFF", k, "max abs %.3e" % np.nanmax(d), "n differing", int((d > 0).sum()), "of", d.size, " above 1e-6:", big)
    print("compared %d arrays, %d differ" % (len(A.files), bad))
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "compare":
        sys.exit(1 if compare(sys.argv[2], sys.argv[3]) else 0)
    run(sys.argv[1])
