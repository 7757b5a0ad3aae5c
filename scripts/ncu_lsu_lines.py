"""Shared-memory wavefronts and global sectors per CUDA source line of one kernel of an ncu report.
usage: python scripts/ncu_lsu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys, os, collections
rep, kre = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {k: j for j, k in enumerate(r)}; H = r; continue
    if hdr is None or len(r) < len(H): continue
    if r[0] == "" or r[2] != "-": continue
    def f(k):
        try: return float(r[hdr[k]] or 0)
        except (ValueError, KeyError): return 0.0
    lines.append((fname, int(r[0]), r[1].strip()[:80], f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal"), f("L2 Theoretical Sectors Global"),
                  f("Instructions Executed"), f("# Samples"), f("L1 Tag Requests Global")))
tw = sum(l[3] for l in lines); tg = sum(l[5] for l in lines); ti = sum(l[6] for l in lines); ts = sum(l[7] for l in lines); tt = sum(l[8] for l in lines)
print("shared wavefronts %.3e (ideal %.3e), global sectors %.3e, global tag requests %.3e, warp instructions %.3e" % (tw, sum(l[4] for l in lines), tg, tt, ti))
print("--- top lines by shared wavefronts + global tag requests ---")
for l in sorted(lines, key=lambda x: -(x[3] + x[8]))[:top]:
    print("%5.2f%% shw (x%.2f of ideal) %5.2f%% gtag %5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * l[3] / max(tw, 1), l[3] / max(l[4], 1), 100 * l[8] / max(tt, 1), 100 * l[6] / ti,
                                                                       100 * l[7] / max(ts, 1), l[0][:12], l[1], l[2]))
