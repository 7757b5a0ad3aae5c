"""Parity checks of the CUDA path against the CPU oracle -- TEST INFRASTRUCTURE (the checker, never the product):
used by tests/ and by bench.py's post-timing `parity_sample` leg only."""
import numpy as np

from . import oracle as O


def oracle_sensitivity(model, sdf, st, pr, idx, ref, tol, amplitude=1e-12, seeds=(101, 102, 103, 104), nthreads=8):
    """For problems `idx`: does the ORACLE's own result move by more than `tol` (or change its iteration count) when
    the initial trajectory moves by `amplitude` rad?  Such a problem sits on a branch (LM accept/reject at the
    fidelity threshold, the 1e-5 absolute-decrease stop, a hinge or SDF cell face) that rounding decides; two correct
    fp64 implementations with a different operation order cannot be expected to agree on it.  -> bool mask over idx."""
    sens = np.zeros(len(idx), dtype=bool)
    if len(idx) == 0:
        return sens
    sub = {k: np.ascontiguousarray(v[idx]) for k, v in pr.items()}
    for sd in seeds:
        rng = np.random.default_rng(sd)
        t = sub["init_traj"] + amplitude * rng.standard_normal(sub["init_traj"].shape)
        r2 = O.batch_optimize(model, sdf, sub["start_conf"], sub["start_vel"], sub["end_conf"], sub["end_vel"], t, st,
                              nthreads=nthreads)
        dev = np.abs(r2["traj"] - ref["traj"][idx]).max(axis=1)
        sens |= (dev >= tol) | (r2["iters"] != ref["iters"][idx])
    return sens


def sample_parity(model, sdf, st, pr, got_traj, got_iters, n_sample=512, seed=0, tol=1e-6, nthreads=8):
    """Compare the CUDA results `got_traj` / `got_iters` of the whole batch `pr` with the oracle on a random sample of
    n_sample problems.  -> dict(n, match_frac, max_abs_rad over the matching problems, max over all, mismatch
    classification as in oracle_sensitivity)."""
    B = pr["init_traj"].shape[0]
    idx = np.sort(np.random.default_rng(seed).choice(B, size=min(n_sample, B), replace=False))
    sub = {k: np.ascontiguousarray(v[idx]) for k, v in pr.items()}
    ref = O.batch_optimize(model, sdf, sub["start_conf"], sub["start_vel"], sub["end_conf"], sub["end_vel"],
                           sub["init_traj"], st, nthreads=nthreads)
    d = np.abs(np.asarray(got_traj)[idx] - ref["traj"]).max(axis=1)
    same = (np.asarray(got_iters)[idx] == ref["iters"]) & (d < tol)
    bad = np.nonzero(~same)[0]
    sens = oracle_sensitivity(model, sdf, st, sub, bad, ref, tol, nthreads=nthreads)
    return {"n": int(len(idx)), "match_frac": float(same.mean()),
            "max_abs_rad": float(d[same].max()) if same.any() else None, "max_abs_rad_all": float(d.max()),
            "n_mismatch": int(len(bad)), "n_mismatch_oracle_sensitive": int(sens.sum()),
            "n_mismatch_unexplained": int(len(bad) - sens.sum()), "tolerance_rad": tol,
            "checker": "CPU oracle port (oracle/gpmp2_oracle.cpp), not upstream gpmp2/GTSAM"}
