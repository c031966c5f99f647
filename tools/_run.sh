mkdir -p gpurun_out/r02b
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --train-steps 3 --dump-prof gpurun_out/r02b/prof_a.json > gpurun_out/r02b/bench_a.json 2> gpurun_out/r02b/bench_a.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02b/bench_a.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("train_step",{}).get("ms_per_step"), d["clocks"], d["roofline"]["frac"])
for k,v in d["roofline"]["families"].items():
    if v["ms"]>0.3: print(k, v)
P
