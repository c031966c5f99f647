# Round-end evidence run on one B200 (gpurun): bench line, per-shape event profile, ncu launch list + DRAM traffic pass of the
# same command, the other BASELINE configs, the parity report, smoke, one --set full capture of the gated stencil.
set -x
O=gpurun_out/r02b
mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-prof $O/per_shape_event_profile.json --dump-prof-train $O/per_shape_event_profile_train.json > $O/bench_prof.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --ncu --steps 1 > $O/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/traffic.csv python bench.py --ncu --steps 1 > $O/ncu_traffic.log 2>&1
for c in 2 4 5; do python bench.py --config $c --steps 10 --warmup 3 > $O/bench_cfg$c.json 2>/dev/null; done
timeout 1500 python -m tests.gpu_checks --isolate > $O/gpu_parity_report.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dwconv3x3_tma -s 3 -c 1 -o $O/ncu_dwg512 -f python tools/probe.py dwg512 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:small_ci -s 3 -c 1 -o $O/ncu_sci -f python tools/probe.py sci_b4 > /dev/null 2>&1
tail -3 $O/gpu_parity_report.txt; tail -2 $O/smoke.log; tail -c 600 $O/bench_n1.json
