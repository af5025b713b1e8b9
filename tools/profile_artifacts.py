"""Turn the CSV exports of tools/ncu_capture.sh (gpurun_out/r02_*) into the tracked artefacts under profiles/:
the launch list, the `--set full` summary, traffic.json (stamped with the sha1 of csrc/sgbm.cu) and the SASS extracts."""
import csv
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def traffic(path):
    r = list(csv.reader(open(path)))
    H, U = r[0], r[1]
    ki, rd, wr = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    out = {}
    for x in r[2:]:
        name = x[ki].split("(")[0].split("::")[-1].split("<")[0]
        out.setdefault(name, []).append(float(x[rd]) * scale[U[rd]] + float(x[wr]) * scale[U[wr]])
    return {k: sum(v) / len(v) for k, v in out.items()}


def main(tag="r02"):
    shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches.csv"))
    with open(os.path.join(P, tag + "_ncu_full_summary.txt"), "w") as fh:
        for part in ("a", "b"):
            fh.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "%s_full_%s_raw.csv" % (tag, part))],
                                    capture_output=True, text=True).stdout)
    t = {}
    for part in ("a", "b"):
        t.update(traffic(os.path.join(G, "%s_full_%s_raw.csv" % (tag, part))))
    sha = hashlib.sha1(open(os.path.join(ROOT, "openvo_b200", "csrc", "sgbm.cu"), "rb").read()).hexdigest()[:16]
    json.dump({"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (tools/ncu_capture.sh, profiles/%s_ncu_full_summary.txt), "
                        "24 KITTI frames per launch; bench.py quotes a figure only while source_sha1 matches openvo_b200/csrc/sgbm.cu" % tag,
               "source_sha1": sha,
               "K": {"k_sgbm_cost_t": t["k_sgbm_cost"], "k_sgbm_vsum_t": t["k_sgbm_vsum"], "k_sgbm_horiz_t": t["k_sgbm_horiz"]}},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_extract.py"), tag], check=True)
    print(json.dumps(t, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
