"""Executed warp-instructions and stall samples per code region (source line ranges given as name:file:lo-hi)."""
import csv, subprocess, sys, os
rep = sys.argv[1]
regions = [a.split(":") for a in sys.argv[2:]]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; L = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {k: j for j, k in enumerate(r)}; H = r; continue
    if hdr is None or len(r) < len(H) or r[0] == "" or r[2] != "-": continue
    g = lambda k: float(r[hdr[k]] or 0)
    L.append((fname, int(r[0]), g("# Samples"), g("Instructions Executed")))
ts = sum(x[2] for x in L); ti = sum(x[3] for x in L)
print("total samples %d  warp-instr %.3e" % (ts, ti))
for name, f, rng in regions:
    lo, hi = map(int, rng.split("-"))
    s = sum(x[2] for x in L if x[0].startswith(f) and lo <= x[1] <= hi); i = sum(x[3] for x in L if x[0].startswith(f) and lo <= x[1] <= hi)
    print("%-14s samples %5.1f%%  instr %5.1f%% (%.3e)" % (name, 100 * s / ts, 100 * i / ti, i))
