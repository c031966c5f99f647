set -x
mkdir -p gpurun_out/r02b
python -m pytest tests -m gpu -x -q > gpurun_out/r02b/pytest_pdl.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02b/pytest_pdl.log
TDR_PDL=0 python bench.py --no-cpu-baseline --train-steps 5 > gpurun_out/r02b/bench_pdl0.json 2> gpurun_out/r02b/bench_pdl0.err
python bench.py --no-cpu-baseline --train-steps 5 > gpurun_out/r02b/bench_pdl1.json 2> gpurun_out/r02b/bench_pdl1.err
TDR_PDL=0 python bench.py --no-cpu-baseline --train-steps 0 > gpurun_out/r02b/bench_pdl0b.json 2>/dev/null
python bench.py --no-cpu-baseline --train-steps 0 > gpurun_out/r02b/bench_pdl1b.json 2>/dev/null
python - <<'P'
import json
for n in ("pdl0","pdl1","pdl0b","pdl1b"):
    try:
        d=json.loads(open(f"gpurun_out/r02b/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d.get("train_step",{}).get("ms_per_step"), d["clocks"])
    except Exception as e: print(n, "ERR", e)
P
