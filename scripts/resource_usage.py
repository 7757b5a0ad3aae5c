"""Static evidence of the built library (no GPU needed): registers / stack / static shared memory of the kernels on the
benched paths (cuobjdump --dump-resource-usage) and their SASS mnemonic counts (cuobjdump -sass) -- DMMA proves the FP64
tensor-core solve, STL / LDL show local-memory traffic.  Usage: python scripts/resource_usage.py > profiles/rNN_resource_usage.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpmp2_b200", "csrc", "libgpmp2b.so")
WANT = [("WAM pipeline", "pk_linh_kernel<VecOpt<7, 3, false>"), ("WAM pipeline", "pk_solve_mma_h_kernel<7>"),
        ("WAM pipeline", "pk_err_kernel<VecOpt<7, 3, false>"), ("WAM one-kernel LM", "gpmp2b_kernel<VecOpt<7, 3, false>, 1>"),
        ("config 4 pipeline", "pk_lin_full_kernel<LieOpt<5, 2, false>"), ("config 4 pipeline", "pk_solve_mma_h_kernel<5>"),
        ("config 4 pipeline", "pk_err_kernel<LieOpt<5, 2, false>"), ("config 4 one-kernel LM", "gpmp2b_kernel<LieOpt<5, 2, false>, 1>"),
        ("config 2 (3-link planar)", "gpmp2b_kernel<VecOpt<3, 2, false>, 1>"), ("config 1 (2-link planar)", "gpmp2b_kernel<VecOpt<2, 2, false>, 1>")]
COUNT = ("DMMA", "DFMA", "DMUL", "DADD", "SHFL", "LDS", "STS", "LDL", "STL")


def main():
    ru = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    rows, name = {}, None
    for ln in ru:
        m = re.search(r"Function (\S+):", ln)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and name:
            rows[name] = tuple(int(x) for x in m.groups())
            name = None
    dem = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
    short = lambda d: d.split("(")[0].replace("void ", "")[:58]
    print("# cuobjdump --dump-resource-usage gpmp2_b200/csrc/libgpmp2b.so (sm_100a): kernels of the benched paths")
    print("# %-26s %-58s %4s %6s %s" % ("path", "kernel", "REG", "STACK", "SHARED(static)"))
    sel = []
    for path, pat in WANT:
        for mang, d in zip(rows, dem):
            if d.startswith("void " + pat) or d.startswith(pat):
                r = rows[mang]
                print("  %-26s %-58s %4d %6d %7d" % (path, short(d), r[0], r[1], r[2]))
                sel.append((mang, d))
    print("\n# all %d kernels of the library: max REG %d; kernels with a stack frame (indexed local arrays or spills): %d"
          % (len(rows), max(r[0] for r in rows.values()), sum(1 for r in rows.values() if r[1] > 0)))
    print("\n# SASS mnemonic counts (cuobjdump -sass -fun <kernel>), static instruction counts")
    for mang, d in sel:
        s = subprocess.run(["cuobjdump", "-sass", "-fun", mang, LIB], capture_output=True, text=True).stdout
        c = collections.Counter()
        for ln in s.splitlines():
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            op = m.group(1)
            base = op.split(".")[0]
            c["total"] += 1
            if base in COUNT:
                c[base] += 1
            if base == "LDG":
                c[op] += 1
        ldg = " ".join("%s %d" % kv for kv in sorted(c.items()) if kv[0].startswith("LDG"))
        print("  %-58s total %6d  %s  %s" % (short(d), c["total"], "  ".join("%s %d" % (k, c[k]) for k in COUNT), ldg))


if __name__ == "__main__":
    main()
