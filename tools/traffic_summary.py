"""Summarise `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` of
`bench.py --ncu --steps 1` (profiler range = one steady-state forward step) into profiles/.

    python tools/traffic_summary.py gpurun_out/traffic.csv 512x512_b4 [r01f]      (last argument: round tag of the .txt)
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("conv_gemm_kernel", "tdr_conv_gemm"), ("dwconv3x3", "tdr_dwconv3x3"), ("rownorm", "tdr_rownorm"),
          ("mdta_gram", "tdr_mdta_gram"), ("mdta_softmax", "tdr_mdta_weff"), ("mdta_fold", "tdr_mdta_weff"),
          ("transfer_kernel", "tdr_masa_transfer"), ("conv3x3_small_ci", "tdr_conv3x3_small_ci")]
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NS = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


TAG = sys.argv[3] if len(sys.argv) > 3 else "r01f"


def main(path, workload):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    per = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
        d = per.setdefault(int(r[0]), dict(name=name))
        d[r[12]] = (float(r[14]), r[13])
    fam = collections.defaultdict(lambda: dict(n=0, t=0.0, rd=0.0, wr=0.0))
    for d in per.values():
        f = next((v for k, v in FAMILY if d["name"].startswith(k)), d["name"][:40])
        a = fam[f]
        a["n"] += 1
        v, u = d["gpu__time_duration.sum"]; a["t"] += v * NS.get(u, 1)
        v, u = d["dram__bytes_read.sum"]; a["rd"] += v * BYTES[u]
        v, u = d["dram__bytes_write.sum"]; a["wr"] += v * BYTES[u]
    out = {}
    txt = os.path.join(ROOT, "profiles", f"{TAG}_dram_traffic_forward_step.txt")
    with open(txt, "w") as fh:
        fh.write("ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                 "--clock-control none  python bench.py --ncu --steps 1\n")
        fh.write(f"one steady-state forward step of RestormerRefFusion option-003, workload {workload}; per kernel family, "
                 "summed over its launches\n")
        fh.write(f"{'family':<34}{'n':>5}{'ms':>9}{'DRAM rd GB':>12}{'DRAM wr GB':>12}{'GB/launch':>11}\n")
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["t"]):
            fh.write(f"{k:<34}{v['n']:>5}{v['t'] / 1e6:>9.3f}{v['rd'] / 1e9:>12.3f}{v['wr'] / 1e9:>12.3f}"
                     f"{(v['rd'] + v['wr']) / 1e9 / v['n']:>11.4f}\n")
            out[k] = dict(launches=v["n"], dram_bytes_per_launch=(v["rd"] + v["wr"]) / v["n"], dram_bytes=v["rd"] + v["wr"],
                          ms=v["t"] / 1e6)
        tot = sum(v["rd"] + v["wr"] for v in fam.values())
        fh.write(f"total DRAM traffic of the step: {tot / 1e9:.2f} GB in {sum(v['n'] for v in fam.values())} launches\n")
    json.dump(dict(source=f"profiles/{TAG}_dram_traffic_forward_step.txt (ncu metrics pass on B200)", workload=workload,
                   families=out), open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    print(open(txt).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "512x512_b4")
