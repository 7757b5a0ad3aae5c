"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {k: j for j, k in enumerate(r)}; H = r; continue
    if hdr is None or len(r) < len(H): continue
    if r[0] == "" or r[2] != "-": continue   # sass rows
    def f(k):
        try: return float(r[hdr[k]] or 0)
        except ValueError: return 0.0
    stalls = {k[6:]: f(k) for k in H if k.startswith("stall_") and "Not Issued" not in k}
    lines.append((fname, int(r[0]), r[1].strip()[:90], f("# Samples"), f("Instructions Executed"), stalls))
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines)
print("total samples %d, total warp instructions %.3e" % (ts, ti))
agg = {}
for l in lines:
    for k, v in l[5].items(): agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(sum(agg.values()), 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
by_file = {}
for l in lines:
    a = by_file.setdefault(l[0], [0, 0]); a[0] += l[3]; a[1] += l[4]
for k, v in by_file.items(): print("  %-28s samples %5.1f%%  inst %5.1f%%" % (k, 100 * v[0] / ts, 100 * v[1] / ti))
print("--- top lines by samples ---")
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    st = sorted(l[5].items(), key=lambda x: -x[1])[:2]
    print("%5.2f%% s %5.2f%% i  %s:%d  [%s]  %s" % (100 * l[3] / ts, 100 * l[4] / ti, l[0][:14], l[1],
          ",".join("%s %.0f%%" % (k, 100 * v / max(l[3], 1)) for k, v in st), l[2]))
