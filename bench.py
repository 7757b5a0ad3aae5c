#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched GPMP2 trajectory-optimization hot path.

  python bench.py --gpus N --steps K --warmup W           (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU reference arm: the oracle, all host threads)

A "step" = one pass of the hot path over one batch: B WAM 7-DOF problems (10 support intervals,
300^3 SDF, obs_check_inter 5, LM, 10 iterations) optimized by ONE gpmp2b_batch_optimize call per rank.
  value : whole-job trajectories/s, inputs resident in HBM (device pointers), CUDA-event timed
  e2e   : the same through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "trajectories/s"
FP64_SPEC_TFLOPS = 37.2                                            # 148 SM x 64 FMA/clk x 2 x 1.965 GHz
SWEEP_BATCHES = (1024, 4096, 16384, 65536, 262144, 1048576)        # config 5
METRICS = {
    "wam": "WAM 7-DOF trajectories optimized/sec",
    "sweep": "WAM 7-DOF trajectories optimized/sec",
    "planar2": "planar 2-link trajectories optimized/sec",
    "planar3gp": "planar 3-link (GP-interpolated obstacle factors) trajectories optimized/sec",
    "mobile": "Pose2MobileArm trajectories optimized/sec",
}
CPU_KIND_NOTE = "oracle port (CPU restatement of the reference path; NOT upstream gpmp2/GTSAM, which cannot be built here)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="wam", choices=["wam", "planar2", "planar3gp", "mobile", "sweep"],
                    help="BASELINE.json config (default: the headline WAM workload, configs[2])")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch problems per GPU; strong: --batch problems in total, sharded over the GPUs")
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU per step (strong scaling: in total); 0 = the config's")
    ap.add_argument("--sdf", type=int, default=300, help="SDF cells per axis (WAM configs)")
    ap.add_argument("--mode", default="restart", choices=["restart", "random"])
    ap.add_argument("--inter", type=int, default=-1, help="obs_check_inter (-1 = the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-sample", action="store_true")
    ap.add_argument("--parity-sample", type=int, default=512, help="problems of the last timed step checked against the oracle")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def load_config(args):
    from gpmp2_b200 import synth
    cfg = synth.baseline_config(args.config, sdf_cells=args.sdf, inter=None if args.inter < 0 else args.inter)
    cfg["work"] = synth.algorithmic_work(cfg)
    if not args.batch:
        args.batch = cfg["batch"]
    return cfg


def workload(args, cfg, B_local, world):
    st = cfg["setting"]
    w = {
        "workload": "%s, total_step=%d (%d support states), obs_check_inter=%d, LM lambda0=100, max_iter=%d, rel_thresh=0, "
                    "%s problems, batch %d per GPU per step (%s scaling, %d in total)"
                    % (cfg["label"], st.total_step, st.total_step + 1, cfg["K"], st.max_iter,
                       ("random-restart" if args.mode == "restart" else "random start/goal") if cfg["name"] in ("wam", "sweep") else "random start/goal",
                       B_local, args.scaling, B_local * world),
        "baseline_config": cfg["name"], "batch_per_gpu": B_local, "batch_total": B_local * world,
        "obs_check_inter": cfg["K"], "max_iter": st.max_iter, "dof": cfg["D"], "spheres": cfg["S"],
    }
    if cfg["ndim"] == 3:
        w["sdf_cells"] = args.sdf
        w["l2"] = ("inputs larger than L2 (%.0f MB SDF in quad cells + %.0f MB trajectories per step, a different seeded "
                   "problem set each step)" % (args.sdf ** 3 * 32 / 1e6, B_local * 2 * cfg["N"] * cfg["D"] * 8 / 1e6))
    else:
        w["l2"] = ("L2 flushed between timed steps (a 256 MB buffer is rewritten; the 2-D field itself is %.1f MB in quad "
                   "cells and would otherwise stay L2-resident); a different seeded problem set each step"
                   % (cfg["sdf"]._rows * cfg["sdf"]._cols * 32 / 1e6))
    return w


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(args, cfg, step, rank, B):
    """The problem set of one step: a different seeded set per (rank, step)."""
    return cfg["problems"](B, 1000 * rank + step + cfg["seed"], args.mode)


def cpu_baseline(args, cfg, nthreads, sample):
    from oracle import oracle as O
    pr = make_inputs(args, cfg, 0, 0, sample)
    t0 = time.perf_counter()
    O.batch_optimize(cfg["model"], cfg["sdf"], pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"],
                     pr["init_traj"], cfg["setting"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, dt


def cpu_sample_size(args, cfg, cores):
    if args.cpu_sample:
        return args.cpu_sample
    # ~15-25 s of CPU work: ~7 ms per WAM problem per core, scaled by the algorithmic work of the config
    w = cfg["work"]
    rel = (w["mflop_linearize"] + 1.5 * w["mflop_solve"] + 1.6 * w["mflop_error_eval"]) / 0.57
    return int(cores * 1536 / max(rel, 0.05))


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path -- here the oracle port (gpmp2 + GTSAM
    cannot be built in this image), one problem per host thread, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    cfg = load_config(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B_local = args.batch // world if args.scaling == "strong" else args.batch
    sample = cpu_sample_size(args, cfg, cores)
    if not args.cpu_sample and args.steps > 14:
        # keep the whole --steps K run within a few minutes (~150 s of CPU work): the per-step sample shrinks with K
        sample = max(cores * 64, sample * 14 // args.steps)
    for s in range(args.warmup):
        cpu_baseline(args, cfg, cores, max(cores, 8))
    t_tot = 0.0
    for s in range(args.steps):
        _, dt = cpu_baseline(args, cfg, cores, sample)
        t_tot += dt
    value = sample * args.steps / t_tot
    line = {
        "impl": "reference", "metric": METRICS[cfg["name"]], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args, cfg, B_local, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d problems of the same workload per step, one problem per thread on %d threads; %s"
                                   % (sample, cores, CPU_KIND_NOTE)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:   # a real gpmp2 + GTSAM build, if one is ever importable, is diffed against the oracle (oracle/probe_reference.py)
        from oracle import probe_reference
        pr = probe_reference.probe(write=False)
        line["reference_probe"] = {k: pr[k] for k in pr if k != "tried"}
    except Exception as e:   # noqa: BLE001
        line["reference_probe"] = {"found": False, "error": str(e)}
    emit(line)


_JSON_FD = None


def _claim_stdout():
    """Libraries (NCCL prints its version banner there) must not share stdout with the ONE JSON line: point fd 1 at
    stderr for everything else and keep the real stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def kernel_source_sha():
    """Identity of the kernel sources: an ncu capture only describes the kernel it was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gpmp2_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cuh", ".cu", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measure_peaks(ctx):
    """Roofline denominators measured live (MEASURED_PEAKS.json has only HBM + bf16): three runs, all recorded, the
    MAX of each is the stated peak (round 1 saw a 12 % run-to-run swing of the L2 figure)."""
    runs = [ctx.measure_peaks() for _ in range(3)]
    best = {k: max(r[k] for r in runs) for k in runs[0]}
    return best, runs


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import gpmp2_b200 as G
    from gpmp2_b200.distributed import ShardedPlanner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = G.Context(local)
    cfg = load_config(args)                  # every rank builds the same seeded scene: zero setup communication
    model, sdf, st = cfg["model"], cfg["sdf"], cfg["setting"]
    D, N = cfg["D"], cfg["N"]
    TL = 2 * N * D
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg["ndim"] == 2 else None   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_case(B, steps, warmup, with_e2e=True):
        """Time `steps` steps of B problems per GPU.  -> dict (rank-0 view; times are max over ranks)."""
        nsteps = warmup + steps
        nsets = nsteps if B <= 65536 else 2          # distinct problem sets per step (two alternating ones for huge batches)
        hsets = [make_inputs(args, cfg, s, rank, B) for s in range(nsets)]
        dsets = [{k: torch.from_numpy(v).to(dev) for k, v in pr.items()} for pr in hsets]
        pl = ShardedPlanner(ctx, model, sdf, st, B, dev)
        for s in range(warmup):
            pl.run_device(dsets[s % nsets])
        barrier()
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kern_ms = 0.0
        barrier()
        if flush is None:
            evs[0][0].record(stream)
            for s in range(warmup, nsteps):
                pl.run_device(dsets[s % nsets])
            evs[0][1].record(stream)
            barrier()
            total_ms = evs[0][0].elapsed_time(evs[0][1])
        else:
            for i, s in enumerate(range(warmup, nsteps)):
                flush.add_(1)                        # rewrite 256 MB: evicts the field and the previous step's data from L2
                evs[i][0].record(stream)
                pl.run_device(dsets[s % nsets])
                evs[i][1].record(stream)
            barrier()
            total_ms = sum(a.elapsed_time(b) for a, b in evs)
        clocks = sampler.stop() if rank == 0 else None
        gpu_launches = ctx.launch_count() - launches0
        ks = ctx.last_kernel_stats()                 # kernel-only stats of the last timed step (library events)
        iters_mean = float(pl.ints[0].double().mean().item())
        tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
        out = {"B": B, "total_ms": total_ms, "value": world * B * steps / (total_ms * 1e-3), "ks": ks, "clocks": clocks,
               "gpu_launches": gpu_launches, "launches_per_step": gpu_launches // max(steps, 1), "iters_mean": iters_mean, "last_inputs": hsets[(nsteps - 1) % nsets],
               # snapshots: at N > 1 the end-to-end leg below reuses the planner's result buffers
               "last_traj": pl.traj.clone(), "last_iters": pl.ints[0].clone()}
        if with_e2e:
            out.update(run_e2e(B, steps, hsets, pl))
        return out

    def run_e2e(B, steps, hsets, pl):
        """The same metric end to end: pinned HOST inputs, H2D + D2H (and at N > 1 the gather to rank 0) inside the timed
        region.  N = 1: one gpmp2b_batch_optimize call with host buffers (the C ABI a user binds).  N > 1: the sharded
        user call (ShardedPlanner.run_host): H2D of the shard, optimize, gather to rank 0, D2H of everything there."""
        def pinned(a):
            t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype).pin_memory()
            t.numpy()[...] = a
            return t
        hp = [{k: pinned(v) for k, v in hsets[s].items()} for s in range(min(len(hsets), 2))]
        if world == 1:
            import ctypes as C
            h_out = torch.empty((B, TL), dtype=torch.float64).pin_memory()
            h_err, h_cc = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory()
            h_it, h_st = torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()
            sset, keep = st.pack()

            def step_host(s):
                h = hp[s % len(hp)]
                ctx.check(ctx.lib.gpmp2b_batch_optimize(
                    ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(sset), B,
                    h["start_conf"].data_ptr(), h["start_vel"].data_ptr(), h["end_conf"].data_ptr(), h["end_vel"].data_ptr(),
                    h["init_traj"].data_ptr(), h_out.data_ptr(), h_err.data_ptr(), h_cc.data_ptr(), h_it.data_ptr(),
                    h_st.data_ptr(), 0, None))
            d2h = B * (TL * 8 + 8 + 8 + 4 + 4)
        else:
            def step_host(s):
                pl.run_host(hp[s % len(hp)])
                torch.cuda.synchronize()
            d2h = pl.d2h_bytes_root()              # on rank 0: the gathered results of all ranks
        step_host(0)
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            if flush is not None:
                flush.add_(1)
            step_host(s)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return {"e2e_value": world * B * steps / float(te.item()), "h2d": B * (4 * D + TL) * 8, "d2h": d2h}

    def roofline(res, peaks, peak_runs):
        ks, w = res["ks"], cfg["work"]
        mflop = ks["linearizations"] * w["mflop_linearize"] + ks["solves"] * w["mflop_solve"] + ks["error_evals"] * w["mflop_error_eval"]
        l2_bytes = (ks["linearizations"] + ks["error_evals"]) * w["lookups_per_pass"] * w["l2_bytes_per_lookup"]
        t_k = ks["kernel_ms"] * 1e-3
        fp64_achieved = mflop * 1e6 / t_k / 1e12
        l2_achieved = l2_bytes / t_k / 1e9
        fp64_frac = fp64_achieved / peaks["fp64_tflops"]
        l2_frac = l2_achieved / peaks["l2_gather_sector_gbs"]
        traffic = None   # DRAM bytes per launch: only from an ncu --set full capture of THIS kernel source and workload
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            for ent in tj.get("captures", []):
                if (ent["kernel_source_sha"] == kernel_source_sha() and ent["baseline_config"] == cfg["name"]
                        and ent["batch_per_gpu"] == res["B"]):
                    traffic = ent["dram_bytes_per_launch"]
        except Exception:
            pass
        # which kernels ran: one fused optimizer kernel per call / chunk, or the phase-kernel pipeline (~70 launches per chunk)
        fused = res.get("launches_per_step", 0) <= 16
        kname = ("gpmp2b_kernel<%s<%d,%d>,LM>" % ("LieOpt" if cfg["lie"] else "VecOpt", D, cfg["ndim"]) if fused else
                 "%s<%s<%d,%d>> + pk_solve_mma_h_kernel<%d> + pk_err_kernel<%s<%d,%d>> (phase-kernel pipeline, %d launches per step)"
                 % ("pk_lin_full_kernel" if cfg["lie"] else "pk_linh_kernel", "LieOpt" if cfg["lie"] else "VecOpt", D, cfg["ndim"], D,
                    "LieOpt" if cfg["lie"] else "VecOpt", D, cfg["ndim"], res.get("launches_per_step", 0)))
        # achieved / peak / unit are those of the binding resource, so that frac = achieved / peak; both resources in full below
        fp64_binds = fp64_frac >= l2_frac
        return {
            "bound": "fp64" if fp64_binds else "l2",
            "achieved": fp64_achieved if fp64_binds else l2_achieved,
            "peak": peaks["fp64_tflops"] if fp64_binds else peaks["l2_gather_sector_gbs"],
            "unit": "TFLOP/s" if fp64_binds else "GB/s",
            "frac": max(fp64_frac, l2_frac), "traffic": traffic, "kernel": kname, "kernel_ms": ks["kernel_ms"],
            "fp64": {"achieved_tflops": fp64_achieved, "peak_tflops_measured": peaks["fp64_tflops"],
                     "peak_tflops_spec": FP64_SPEC_TFLOPS, "frac": fp64_frac},
            "l2": {"achieved_gbs": l2_achieved, "peak_gbs_measured_32B_gather": peaks["l2_gather_sector_gbs"],
                   "peak_gbs_measured_8B_gather": peaks["l2_gather_useful_gbs"], "frac": l2_frac},
            "peak_runs": peak_runs,
            "hbm_algorithmic_gbs": (res.get("h2d", 0) + res.get("d2h", 0)) / t_k / 1e9,
            "per_launch": {"linearizations": ks["linearizations"], "solves": ks["solves"], "error_evals": ks["error_evals"],
                           "algorithmic_mflop": mflop, "algorithmic_l2_bytes": l2_bytes,
                           "mflop_per_linearization": w["mflop_linearize"], "mflop_per_solve": w["mflop_solve"],
                           "mflop_per_error_eval": w["mflop_error_eval"], "l2_bytes_per_lookup": w["l2_bytes_per_lookup"],
                           "lookups_per_pass": w["lookups_per_pass"]},
            "note": "units per launch counted by the kernel x SURVEY.md 8(d) per-unit figures recounted for this config; peak = "
                    "dependent-free DFMA loop / L2-resident random 32-byte gather (the SDF quad-cell access pattern) measured "
                    "live by gpmp2b_measure_peaks, max of 3 runs (MEASURED_PEAKS.json has only HBM + bf16); frac = max(fp64, l2)",
        }

    B_local = max(1, args.batch // world) if args.scaling == "strong" else args.batch
    sweep = None
    if args.config == "sweep":
        # config 5: the WAM inputs at B in {1k .. 1M} (per GPU for weak scaling, in total for strong scaling)
        sweep = []
        for Bs in SWEEP_BATCHES:
            Bl = max(1, Bs // world) if args.scaling == "strong" else Bs
            steps = max(1, min(args.steps, (4 * 65536) // Bl)) if Bl > 65536 else args.steps
            r = run_case(Bl, steps, min(args.warmup, 3) if Bl <= 65536 else 1, with_e2e=False)
            ks = r["ks"]
            sweep.append({"batch_total": Bl * world, "batch_per_gpu": Bl, "steps": steps, "value": r["value"],
                          "ms_per_step": r["total_ms"] / steps, "kernel_ms": ks["kernel_ms"],
                          "us_per_lm_iteration_per_batch": 1e3 * ks["kernel_ms"] / max(r["iters_mean"], 1e-9),
                          "linearizations": ks["linearizations"], "solves": ks["solves"], "error_evals": ks["error_evals"]})
            del r
            torch.cuda.empty_cache()
        B_local = max(1, 65536 // world) if args.scaling == "strong" else 65536
    res = run_case(B_local, args.steps, args.warmup)

    if rank == 0:
        peaks, peak_runs = measure_peaks(ctx)
        ks = res["ks"]
        line = {
            "metric": METRICS[cfg["name"]], "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["total_ms"] / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload(args, cfg, B_local, world),
            "us_per_lm_iteration_per_batch": 1e3 * ks["kernel_ms"] / max(res["iters_mean"], 1e-9),
            "ns_per_lm_iteration_per_trajectory": 1e6 * ks["kernel_ms"] / max(ks["linearizations"], 1),
            "mean_lm_iterations": res["iters_mean"],
            "clocks": res["clocks"], "gpu_launches": res["gpu_launches"],
            "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "path": "gpmp2b_batch_optimize with pinned host buffers" if world == 1 else
                            "ShardedPlanner.run_host: pinned H2D of each shard, optimize, NCCL gather to rank 0, D2H of all results there"},
            "roofline": roofline(res, peaks, peak_runs),
        }
        if sweep is not None:
            line["sweep"] = sweep
        if not args.no_parity_sample:
            # the checker, after the timed region: a sample of the LAST timed step's results against the CPU oracle
            from oracle import oracle as O
            from oracle.parity import sample_parity
            O.build()
            line["parity_sample"] = sample_parity(model, sdf, st, res["last_inputs"], res["last_traj"].cpu().numpy(),
                                                  res["last_iters"].cpu().numpy(), n_sample=args.parity_sample,
                                                  nthreads=os.cpu_count() or 1)
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            cores = os.cpu_count() or 1
            sample = cpu_sample_size(args, cfg, cores)
            v, dt = cpu_baseline(args, cfg, cores, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d problems of the same workload, one problem per thread on %d threads, %.1f s; %s"
                                              % (sample, cores, dt, CPU_KIND_NOTE)}
        emit(line)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
