"""Static SASS size of one kernel per source region (nvdisasm --print-line-info on the extracted cubin).
usage: python scripts/code_size.py obj_or_so kernel_substring [file:lo-hi:name ...]"""
import subprocess, sys, re, os, tempfile, glob, collections
obj, pat = sys.argv[1], sys.argv[2]
regs = []
for a in sys.argv[3:]:
    f, rng, name = a.split(":"); lo, hi = map(int, rng.split("-")); regs.append((f, lo, hi, name))
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
cnt = collections.Counter(); lines = collections.Counter()
for cub in glob.glob(td + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    on = False; cur = None
    for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m: on = pat in m.group(1); continue
        if l.strip().startswith(".section"): on = False
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l) and cur:
            for f, lo, hi, name in regs:
                if cur[0].startswith(f) and lo <= cur[1] <= hi: cnt[name] += 1; break
            else: cnt[cur[0]] += 1
            lines[cur] += 1
tot = sum(cnt.values())
for k, v in cnt.most_common(): print("%-24s %5d instr %6.1f KB" % (k, v, v * 16 / 1024))
print("total %d instr %.1f KB" % (tot, tot * 16 / 1024))
if os.environ.get("TOPLINES"):
    for (f, ln), v in lines.most_common(int(os.environ["TOPLINES"])): print("  %s:%d %d" % (f, ln, v))
