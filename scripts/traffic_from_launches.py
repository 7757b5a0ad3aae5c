"""DRAM bytes of ONE optimize step of the phase-kernel pipeline from an ncu launch list (gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum per launch) of `bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-sample`:
the second group of 1 + 3 x rounds pipeline launches is the timed device-buffer step at the full batch.  Writes the entry
bench.py copies into roofline.traffic (keyed by the hash of the kernel sources) into profiles/traffic.json.
usage: python scripts/traffic_from_launches.py launches.csv CONFIG BATCH [launches_per_step=70] [source label]"""
import collections, csv, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import kernel_source_sha  # noqa: E402
path, config, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
per_step = int(sys.argv[4]) if len(sys.argv) > 4 else 70
label = sys.argv[5] if len(sys.argv) > 5 else path
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ii, ki, mi, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
launches = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    else:
        v *= {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1e-6)
    launches.setdefault(r[ii], {"name": r[ki].split("<")[0].replace("void ", "")})[r[mi]] = v
pk = [l for l in launches.values() if l["name"].startswith("pk_")]
assert len(pk) >= 2 * per_step, "expected at least two pipeline steps in the launch list, got %d launches" % len(pk)
step = pk[per_step:2 * per_step]
per = collections.defaultdict(lambda: [0.0, 0.0])
for l in step:
    per[l["name"]][0] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    per[l["name"]][1] += l.get("gpu__time_duration.sum", 0.0)
total = sum(v[0] for v in per.values())
ent = {"kernel_source_sha": kernel_source_sha(), "baseline_config": config, "batch_per_gpu": batch, "dram_bytes_per_launch": int(total),
       "what": "one optimize step = %d launches of the phase-kernel pipeline, device buffers: sum of dram__bytes_read.sum + dram__bytes_write.sum over them" % per_step,
       "per_kernel": {k: {"dram_gb": round(v[0] / 1e9, 3), "ms_under_ncu": round(v[1], 2)} for k, v in sorted(per.items())},
       "source": label}
tj_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
tj = json.load(open(tj_path))
tj["captures"] = [c for c in tj["captures"] if not (c["baseline_config"] == config and c["batch_per_gpu"] == batch and c["kernel_source_sha"] == ent["kernel_source_sha"])] + [ent]
json.dump(tj, open(tj_path, "w"), indent=1)
print(json.dumps(ent, indent=1))
