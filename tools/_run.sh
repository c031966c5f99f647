mkdir -p gpurun_out/r02b
python bench.py --no-cpu-baseline --train-steps 3 > gpurun_out/r02b/bench_c3.json 2> gpurun_out/r02b/bench_c3.err; tail -c 300 gpurun_out/r02b/bench_c3.err
for c in 2 4 5; do
python bench.py --config $c --no-cpu-baseline --train-steps 0 > gpurun_out/r02b/bench_c$c.json 2> gpurun_out/r02b/bench_c$c.err; tail -c 300 gpurun_out/r02b/bench_c$c.err
done
python - <<'P'
import json
for c in (3,2,4,5):
    d=json.loads(open(f"gpurun_out/r02b/bench_c{c}.json").read().strip().splitlines()[-1])
    print(c, round(d["ms_per_step"],3), round(d["value"],2), round(d["e2e"]["value"],2), d.get("eager_launch"), d["config"]["launch"][:60], d["gpu_launches"], d["clocks"]["sm_mhz"], d.get("train_step",{}).get("ms_per_step"))
P
