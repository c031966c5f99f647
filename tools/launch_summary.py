"""Per-kernel launch summary of an ncu CSV (`--metrics gpu__time_duration.sum[,...] --csv`) of `bench.py --ncu --steps 1`.

    python tools/launch_summary.py gpurun_out/traffic.csv > profiles/<round>_launches_step_summary.txt
Times under ncu are cold-cache and serialised: the SHARES are what must agree with bench.py's live event profile.
"""
import collections
import csv
import re
import sys

NS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    per = {}
    for r in rows:
        if r[12] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
        per[int(r[0])] = (name, float(r[14].replace(",", "")) * NS[r[13]])
    fam = collections.defaultdict(lambda: [0, 0.0])
    for name, ms in per.values():
        fam[name][0] += 1
        fam[name][1] += ms
    tot = sum(v[1] for v in fam.values())
    print(f"launches in step: {len(per)}, sum of kernel time {tot:.2f} ms (cold-cache, serialised under ncu)")
    print(f"{'kernel':64s} {'n':>4s} {'ms':>9s} {'share':>7s}")
    for k, (n, ms) in sorted(fam.items(), key=lambda t: -t[1][1]):
        print(f"{k[:64]:64s} {n:4d} {ms:9.3f} {ms / tot:7.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
