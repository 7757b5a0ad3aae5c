"""Raw-metric summary + hottest source lines of one kernel of an ncu --set full report, in the format of profiles/*_ncu_raw_summary.txt.
usage: python scripts/ncu_summary.py report.ncu-rep kernel_regex "header line (the command that produced the report)" > profiles/xyz.txt"""
import csv, subprocess, sys, os
rep, kre, header = sys.argv[1], sys.argv[2], sys.argv[3]
print("# " + header)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]
for k in want:
    if k in h:
        i = h.index(k)
        print("%s = %s %s" % (k, r[i], units[i]))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {}; H = r; [hdr.setdefault(k, j) for j, k in enumerate(r)]; continue
    if hdr is None or len(r) < len(H): continue
    if r[0] == "" or r[2] != "-": continue
    def f(k):
        try: return float(r[hdr[k]] or 0)
        except (ValueError, KeyError): return 0.0
    st = {k[6:]: f(k) for k in hdr if k.startswith("stall_") and "(" not in k}
    lines.append((fname, int(r[0]), r[1].strip()[:100], f("# Samples"), f("Instructions Executed"), st))
ts = sum(l[3] for l in lines) or 1; ti = sum(l[4] for l in lines) or 1
agg = {}
for l in lines:
    for k, v in l[5].items(): agg[k] = agg.get(k, 0) + v
print("total samples %d, source-attributed warp instructions %.3e" % (ts, ti))
print("stall mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / ts) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
byf = {}
for l in lines:
    a = byf.setdefault(l[0], [0, 0]); a[0] += l[3]; a[1] += l[4]
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][0]): print("  %-28s samples %5.1f%%  inst %5.1f%%" % (k, 100 * v[0] / ts, 100 * v[1] / ti))
print("--- top lines by samples ---")
for l in sorted(lines, key=lambda x: -x[3])[:30]:
    s2 = sorted(l[5].items(), key=lambda x: -x[1])[:2]
    print("%5.2f%% s %5.2f%% i  %s:%d  [%s]  %s" % (100 * l[3] / ts, 100 * l[4] / ti, l[0][:16], l[1], ",".join("%s %.0f%%" % (k, 100 * v / max(l[3], 1)) for k, v in s2), l[2]))
