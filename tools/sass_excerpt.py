"""Write profiles/<name>: per-kernel counts of the Blackwell-only SASS mnemonics in the shipped libtdr_sm100.so
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, FFMA2 = packed fp32 FMA) plus the first occurrence of each with its address, from
`cuobjdump -sass`.   python tools/sass_excerpt.py profiles/r02_sass_excerpt.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "textualdegremoval_b200", "libtdr_sm100.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA2", "F2FP.SATFINITE"]


def main(out_path):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    first = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn) + r"\b", line) or (("." in mn) and mn in line):
                per[cur][mn] += 1
                first.setdefault((cur, mn), line.strip()[:150])
    with open(out_path, "w") as fh:
        fh.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]})\n")
        tot = collections.Counter()
        for c in per.values():
            tot.update(c)
        fh.write("# totals: " + "  ".join(f"{k} {tot[k]}" for k in MNEMONICS if tot[k]) + "\n\n")
        fh.write(f"{'kernel':72s} " + " ".join(f"{m:>9s}" for m in MNEMONICS[:9]) + "\n")
        for k, c in per.items():
            if any(c[m] for m in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM")):
                fh.write(f"{k[:72]:72s} " + " ".join(f"{c[m]:9d}" for m in MNEMONICS[:9]) + "\n")
        fh.write("\n# first occurrence per (kernel, mnemonic)\n")
        for (k, mn), line in first.items():
            if mn in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR"):
                fh.write(f"{k[:60]:60s} {line}\n")
    print(open(out_path).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt"))
