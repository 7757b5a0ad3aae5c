"""Print the SASS rows (samples, executed count, top stall) attributed to a CUDA source line range.
usage: python scripts/ncu_sass.py exported_source.csv file_prefix lo hi"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
pref, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
fname = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {k: j for j, k in enumerate(r)}; H = r; continue
    if hdr is None or len(r) < len(H): continue
    if r[0] != "" and r[2] == "-":
        cur = (fname, int(r[0])) ; show = fname.startswith(pref) and lo <= cur[1] <= hi
        if show: print("== %s:%d  %s" % (fname, cur[1], r[1].strip()[:100]))
        continue
    if cur and fname.startswith(pref) and lo <= cur[1] <= hi:
        fl = lambda v: float(v) if v not in ("", "-") else 0.0
        st = {k[6:]: fl(r[hdr[k]]) for k in H if k.startswith("stall_") and "Not Issued" not in k}
        top = sorted(st.items(), key=lambda x: -x[1])[:2]
        print("   %6s smp %10s ex  %-60s %s" % (r[hdr["# Samples"]], r[hdr["Instructions Executed"]], r[3].strip()[:60],
              ",".join("%s %d" % (k, v) for k, v in top if v > 0)))
