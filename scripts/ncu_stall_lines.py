"""Warp-stall samples per CUDA source line of one kernel of an ncu report (source page), with the stall-reason split.
usage: python scripts/ncu_stall_lines.py report.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys, os, collections
rep, kre = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
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
    lines.append((fname, int(r[0]), r[1].strip()[:90], f("# Samples"), f("Instructions Executed"), st))
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines)
tot = collections.Counter()
for l in lines:
    for k, v in l[5].items(): tot[k] += v
print("samples %d, warp instructions %.3e" % (ts, ti))
print("stall reasons: " + ", ".join("%s %.1f%%" % (k, 100 * v / max(ts, 1)) for k, v in tot.most_common(9)))
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    s = ", ".join("%s %.0f%%" % (k, 100 * v / max(l[3], 1)) for k, v in sorted(l[5].items(), key=lambda kv: -kv[1])[:3] if v > 0)
    print("%5.2f%% smp %5.2f%% inst  %s:%d  %-70s [%s]" % (100 * l[3] / max(ts, 1), 100 * l[4] / max(ti, 1), l[0][:14], l[1], l[2][:70], s))
