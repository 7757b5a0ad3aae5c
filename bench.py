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

METRIC = "WAM 7-DOF trajectories optimized/sec"
UNIT = "trajectories/s"
# algorithmic work per unit
MFLOP_LINEARIZE, MFLOP_SOLVE, MFLOP_ERREVAL = 0.32, 0.08, 0.08      # per trajectory
L2_BYTES_PER_LOOKUP = 64.0                                         # 8 doubles per SDF lookup
FP64_SPEC_TFLOPS = 37.2                                            # 148 SM x 64 FMA/clk x 2 x 1.965 GHz


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--sdf", type=int, default=300, help="SDF cells per axis")
    ap.add_argument("--mode", default="restart", choices=["restart", "random"])
    ap.add_argument("--inter", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload(args):
    return {
        "workload": "WAM 7-DOF ArmModel (16 spheres), WAMDeskDataset %d^3 SDF, total_step=10 (11 support states), "
                    "obs_check_inter=%d, LM lambda0=100, max_iter=10, rel_thresh=0, %s problems, batch %d per GPU per step"
                    % (args.sdf, args.inter, "random-restart" if args.mode == "restart" else "random start/goal", args.batch),
        "batch_per_gpu": args.batch, "sdf_cells": args.sdf, "obs_check_inter": args.inter, "max_iter": 10,
        "l2": "inputs larger than L2 (%.0f MB SDF in quad cells + %.0f MB trajectories per step, a different seeded problem set each step)"
              % (args.sdf ** 3 * 32 / 1e6, args.batch * 182 * 8 / 1e6),
    }


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


def make_inputs(args, step, rank, B=None):
    from gpmp2_b200 import synth
    return synth.wam_problems(B or args.batch, total_step=10, seed=1000 * rank + step + 3, mode=args.mode)


def cpu_baseline(args, model, sdf, st, nthreads, sample):
    from oracle import oracle as O
    pr = make_inputs(args, 0, 0, sample)
    t0 = time.perf_counter()
    O.batch_optimize(model, sdf, pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"], pr["init_traj"], st,
                     nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, dt


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path -- here the oracle port (gpmp2 + GTSAM
    cannot be built in this image), one problem per host thread, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpmp2_b200 import synth
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    model = synth.wam_arm()
    sdf = synth.wam_desk_dataset(args.sdf)
    st = synth.bench_setting(7, inter=args.inter)
    sample = args.cpu_sample or cores * 1536   # ~15-25 s of CPU work at ~10 ms per problem per core
    for s in range(args.warmup):
        cpu_baseline(args, model, sdf, st, cores, max(cores, 8))
    t_tot = 0.0
    for s in range(args.steps):
        _, dt = cpu_baseline(args, model, sdf, st, cores, sample)
        t_tot += dt
    value = sample * args.steps / t_tot
    cfg = workload(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d problems of the same workload per step, one problem per thread on %d threads "
                                   "(CPU restatement of the reference path; gpmp2+GTSAM cannot be built here)" % (sample, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import gpmp2_b200 as G
    from gpmp2_b200 import synth
    from gpmp2_b200.distributed import gather_to_root

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
    model = synth.wam_arm()
    sdf = synth.wam_desk_dataset(args.sdf)       # every rank builds the same seeded scene: zero setup comm
    st = synth.bench_setting(7, inter=args.inter)
    B, D, N = args.batch, 7, 11
    TL = 2 * N * D
    nsteps = args.warmup + args.steps

    # ---- device-resident inputs for every step (distinct problem sets) ----
    dsets, hsets = [], []
    for s in range(nsteps):
        pr = make_inputs(args, s, rank)
        hsets.append(pr)
        dsets.append({k: torch.from_numpy(v).to(dev) for k, v in pr.items()})
    out_traj = torch.empty((B, TL), dtype=torch.float64, device=dev)
    out_sc = torch.empty((B, 2), dtype=torch.float64, device=dev)            # error, collision cost
    out_int = torch.empty((B, 2), dtype=torch.int32, device=dev)             # iters, status
    out_err, out_cc = out_sc[:, 0], out_sc[:, 1]
    err_c = torch.empty(B, dtype=torch.float64, device=dev)
    cc_c = torch.empty(B, dtype=torch.float64, device=dev)
    it_c = torch.empty(B, dtype=torch.int32, device=dev)
    stt_c = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step_device(s):
        d = dsets[s]
        G.api.batch_optimize_device(model, sdf, st, B, d["start_conf"].data_ptr(), d["start_vel"].data_ptr(),
                                    d["end_conf"].data_ptr(), d["end_vel"].data_ptr(), d["init_traj"].data_ptr(),
                                    out_traj.data_ptr(), err_c.data_ptr(), cc_c.data_ptr(), it_c.data_ptr(),
                                    stt_c.data_ptr(), stream=stream.cuda_stream, ctx=ctx)
        if world > 1:   # the only communication: final gather of results and costs
            gather_to_root(out_traj)
            gather_to_root(torch.stack([err_c, cc_c], dim=1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        step_device(s)
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ms, lin, sol, evl = 0.0, 0, 0, 0
    barrier()
    ev[0].record(stream)
    for s in range(args.warmup, nsteps):
        step_device(s)
    ev[1].record(stream)
    barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = ctx.launch_count() - launches0
    # kernel-only stats of the last timed step (events recorded by the library on the launching stream)
    ks = ctx.last_kernel_stats()
    iters_mean = float(it_c.double().mean().item())
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- e2e through the C ABI with pinned host buffers ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype).pin_memory()
        t.numpy()[...] = a
        return t

    hp = [{k: pinned(v) for k, v in hsets[s].items()} for s in range(min(nsteps, 2))]
    h_out = torch.empty((B, TL), dtype=torch.float64).pin_memory()
    h_err, h_cc = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory()
    h_it, h_st = torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()
    import ctypes as C
    sset, keep = st.pack()

    def step_host(s):
        h = hp[s % len(hp)]
        ctx.check(ctx.lib.gpmp2b_batch_optimize(
            ctx.h, ctx.robot_handle(model), ctx.sdf_handle(sdf), C.byref(sset), B,
            h["start_conf"].data_ptr(), h["start_vel"].data_ptr(), h["end_conf"].data_ptr(), h["end_vel"].data_ptr(),
            h["init_traj"].data_ptr(), h_out.data_ptr(), h_err.data_ptr(), h_cc.data_ptr(), h_it.data_ptr(),
            h_st.data_ptr(), 0, None))

    step_host(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(s)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())
    h2d = B * (4 * D + TL) * 8
    d2h = B * (TL * 8 + 8 + 8 + 4 + 4)

    if rank == 0:
        peaks = ctx.measure_peaks()
        # roofline of the dominant (only) kernel, per launch = per step per GPU
        mflop = ks["linearizations"] * MFLOP_LINEARIZE + ks["solves"] * MFLOP_SOLVE + ks["error_evals"] * MFLOP_ERREVAL
        lookups = 16 * (10 * (args.inter + 1) + 1)
        l2_bytes = (ks["linearizations"] + ks["error_evals"]) * lookups * L2_BYTES_PER_LOOKUP
        t_k = ks["kernel_ms"] * 1e-3
        fp64_achieved = mflop * 1e6 / t_k / 1e12
        l2_achieved = l2_bytes / t_k / 1e9
        fp64_frac = fp64_achieved / peaks["fp64_tflops"]
        l2_frac = l2_achieved / peaks["l2_gather_sector_gbs"]   # a lookup = two 32-byte quad cells = its 64 algorithmic bytes
        traffic = None   # dram bytes per launch from the committed ncu --set full capture of this exact workload
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            w = tj["workload"]
            if (w["batch_per_gpu"], w["sdf_cells"], w["obs_check_inter"], w["mode"]) == (args.batch, args.sdf, args.inter, args.mode):
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {
            "bound": "fp64", "achieved": fp64_achieved, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
            "frac": max(fp64_frac, l2_frac), "traffic": traffic,
            "kernel": "gpmp2b_kernel<VecOpt<7,3>,LM>", "kernel_ms": ks["kernel_ms"],
            "fp64": {"achieved_tflops": fp64_achieved, "peak_tflops_measured": peaks["fp64_tflops"],
                     "peak_tflops_spec": FP64_SPEC_TFLOPS, "frac": fp64_frac},
            "l2": {"achieved_gbs": l2_achieved, "peak_gbs_measured_32B_gather": peaks["l2_gather_sector_gbs"],
                   "peak_gbs_measured_8B_gather": peaks["l2_gather_useful_gbs"], "frac": l2_frac},
            "hbm_algorithmic_gbs": (h2d + d2h) / t_k / 1e9,
            "per_launch": {"linearizations": ks["linearizations"], "solves": ks["solves"], "error_evals": ks["error_evals"],
                           "algorithmic_mflop": mflop, "algorithmic_l2_bytes": l2_bytes},
            "note": "peak = dependent-free DFMA loop / L2-resident random 32-byte gather (one 256-bit load per lane, the "
                    "SDF quad-cell access pattern) measured live by gpmp2b_measure_peaks (MEASURED_PEAKS.json has only "
                    "HBM + bf16); frac = max(fp64, l2)",
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload(args),
            "us_per_lm_iteration_per_batch": 1e3 * ks["kernel_ms"] / max(iters_mean, 1e-9),
            "ns_per_lm_iteration_per_trajectory": 1e6 * ks["kernel_ms"] / max(ks["linearizations"], 1),
            "mean_lm_iterations": iters_mean,
            "clocks": clocks, "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roof,
        }
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or cores * 1536   # ~15-25 s of CPU work at ~10 ms per problem per core
            v, dt = cpu_baseline(args, model, sdf, st, cores, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d problems of the same workload, one problem per thread on %d threads, %.1f s "
                                              "(CPU restatement of the reference path; gpmp2+GTSAM cannot be built here)"
                                              % (sample, cores, dt)}
        emit(line)
    del keep
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
