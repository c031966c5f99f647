set -x
mkdir -p gpurun_out/r02
python bench.py > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02/per_shape_event_profile.json --dump-prof-train gpurun_out/r02/per_shape_event_profile_train.json > gpurun_out/r02/bench_prof.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02/launches.csv python bench.py --ncu --steps 1 > gpurun_out/r02/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02/traffic.csv python bench.py --ncu --steps 1 > gpurun_out/r02/ncu_traffic.log 2>&1
for c in 2 4 5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02/bench_cfg$c.json 2>/dev/null; done
timeout 1500 python -m tests.gpu_checks --isolate > gpurun_out/r02/gpu_parity_report.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02/smoke.log 2>&1
for p in qkv96 dwg512 vit; do :; done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/r02/ncu_qkv96 -f python tools/probe.py qkv96 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:vit_attn -s 3 -c 1 -o gpurun_out/r02/ncu_vit_attn -f env PYTHONPATH=. python tools/probe_attn.py > /dev/null 2>&1
tail -3 gpurun_out/r02/gpu_parity_report.txt; cat gpurun_out/r02/smoke.log | tail -2; tail -c 400 gpurun_out/r02/bench_n1.json
