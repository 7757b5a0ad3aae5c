"""Executed vs algorithmic FP64 work of the first round of the phase pipeline (scripts/capture_flops.sh).
usage: python scripts/executed_flops.py launches.csv CONFIG [launches.csv CONFIG ...]"""
import collections, csv, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gpmp2_b200 import synth  # noqa: E402
print("# executed FP64 work of round 0 (every trajectory of the batch linearizes / solves / evaluates once) per kernel:")
print("# FLOP = 2 DFMA + DMUL + DADD thread instructions + 512 per DMMA warp instruction (m8n8k4); algorithmic = SURVEY.md 8(d) per unit x batch")
for path, name in zip(sys.argv[1::2], sys.argv[2::2]):
    cfg = synth.baseline_config(name, sdf_cells=10 if name in ("wam", "sweep") else 300)
    w = synth.algorithmic_work(cfg)
    B = cfg["batch"]
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ii, ki, mi, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    L = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            v = float(r[vi].replace(",", ""))
            if r[mi].startswith("gpu__time"):
                v *= {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
            L.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "")})[r[mi].split(".")[0]] = v
    print("config %s (%s), B = %d" % (name, cfg["label"], B))
    for l in L.values():
        kind = "linearize" if "lin" in l["name"] else ("solve" if "solve" in l["name"] else "error_eval")
        alg = w["mflop_" + kind] * 1e6 * B
        fma, mul, add = (l.get("smsp__sass_thread_inst_executed_op_%s_pred_on" % k, 0.0) for k in ("dfma", "dmul", "dadd"))
        dmma = l.get("smsp__inst_executed_pipe_tensor_subpipe_dmma", 0.0)
        ex = 2 * fma + mul + add + 512 * dmma
        ms = l.get("gpu__time_duration", 0.0)
        print("  %-46s %7.3f ms  executed %8.2f GFLOP (DFMA %.3e DMUL %.3e DADD %.3e thread inst, DMMA %.3e warp inst)  algorithmic %7.2f GFLOP  "
              "executed/algorithmic %.2f  executed %.2f TFLOP/s  warp inst %.3e (%.1f k per trajectory)"
              % (l["name"][:46], ms, ex / 1e9, fma, mul, add, dmma, alg / 1e9, ex / alg, ex / ms / 1e9, l.get("smsp__inst_executed", 0.0),
                 l.get("smsp__inst_executed", 0.0) / B / 1e3))
