"""Write compact SASS extracts (address + instruction, opcode histogram) of the main kernels' KITTI-shape instantiations to
profiles/<round>_sass/ — the evidence for the DPX (VIMNMX / VIADDMNMX .U16x2), CREDUX, UBLKCP (TMA bulk copy) and SYNCS (mbarrier)
claims in DESIGN.md.  Usage: python tools/sass_extract.py r02"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"sgbm.cu.o": [("k_sgbm_costILi2ELi16ELb0", "cost_bs5_tx16"), ("k_sgbm_vsumILi8ELi8ELb0ELi16ELi16ELb0", "vsum_d128"),
                      ("k_sgbm_horizILi2ELb0ELi1ELb1ELb0", "horiz_d128")],
        "match.cu.o": [("k_knn2_partial_b", "knn2_partial")]}
KEEP_SUFFIX = ("VIMNMX", "VIADDMNMX", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "LDS", "STS", "LDG", "STG", "SHFL")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, d = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            d[cur].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return d


def main(tag):
    outdir = os.path.join(ROOT, "profiles", tag + "_sass")
    os.makedirs(outdir, exist_ok=True)
    for obj, lst in WANT.items():
        d = functions(os.path.join(ROOT, "openvo_b200", "lib", obj))
        for key, name in lst:
            hit = [k for k in d if key in k][0]
            ops = {}
            for l in d[hit]:
                t = l.split("*/")[1].split()
                op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
                op = op if op.startswith(KEEP_SUFFIX) else op.split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            dem = subprocess.run(["c++filt", hit], capture_output=True, text=True).stdout.strip()
            with open(os.path.join(outdir, name + ".sass"), "w") as fh:
                fh.write("// cuobjdump -sass openvo_b200/lib/%s, function %s (%d instructions)\n" % (obj, dem, len(d[hit])))
                fh.write("// opcode histogram: " + ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:24]) + "\n")
                fh.write("\n".join(d[hit]) + "\n")
            print(name, len(d[hit]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
