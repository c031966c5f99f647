mkdir -p gpurun_out/r02b
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --train-steps 10 --dump-prof-train gpurun_out/r02b/prof_train_b.json > gpurun_out/r02b/bench_b.json 2> gpurun_out/r02b/bench_b.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02b/bench_b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])
t=d["train_step"]; print(t["ms_per_step"], t["value"], t.get("ms_per_step_with_dino_select"))
for k,v in t["roofline"]["families"].items():
    if k in ("tdr_mdta_bwd","tdr_pixel_shuffle_nhwc","tdr_wgrad","tdr_conv_gemm"): print(k, v)
P
