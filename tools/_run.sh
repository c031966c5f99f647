mkdir -p gpurun_out/r02b
TDR_PDL=1 python bench.py --no-cpu-baseline --train-steps 0 > gpurun_out/r02b/bench_gp1.json 2> gpurun_out/r02b/bench_gp1.err; tail -c 300 gpurun_out/r02b/bench_gp1.err
python bench.py --no-cpu-baseline --train-steps 0 > gpurun_out/r02b/bench_gp0.json 2> gpurun_out/r02b/bench_gp0.err
TDR_PDL=1 python bench.py --no-cpu-baseline --train-steps 0 --batch 1 > gpurun_out/r02b/bench_gp1b1.json 2> /dev/null
python bench.py --no-cpu-baseline --train-steps 0 --batch 1 > gpurun_out/r02b/bench_gp0b1.json 2> /dev/null
python - <<'P'
import json
for c in ("gp1","gp0","gp1b1","gp0b1"):
    d=json.loads(open(f"gpurun_out/r02b/bench_{c}.json").read().strip().splitlines()[-1])
    print(c, round(d["ms_per_step"],3), round(d["value"],2), round(d["e2e"]["value"],2), d["eager_launch"]["ms_per_step"], d["config"]["launch"][:40], d["clocks"]["sm_mhz"])
P
