#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_tests.sh r02t
python bench.py --cpu-seconds 8 > gpurun_out/bench_r02t.json 2> gpurun_out/bench_r02t.err; tail -2 gpurun_out/bench_r02t.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02t.json')); r=d['roofline']; e=d['e2e']
print('value %.0f (%.4f ms/step) e2e %.0f frac_h2d %.3f roofline %s frac %.4f alone %s whole %.3f cpu %.0f/%.0f'%(d['value'], d['ms_per_step'], e['value'], e['frac_of_h2d_only'], r['kernel'], r['frac'], r['stage_ms_alone'], r['whole_path']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value']))"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_r02t_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/san_r02t_$tool.txt | tail -3
done
# the registration kernel on the bench workload's first 64 problems under racecheck + synccheck (not just smoke())
for tool in racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_configs.py -x -q -k "bench_workload" > gpurun_out/san_r02t_bench64_$tool.txt 2>&1
  echo "bench64 $tool rc=$?"; grep -E "SUMMARY|passed|failed" gpurun_out/san_r02t_bench64_$tool.txt | tail -3
done
ncu --set full --clock-control none --import-source on -k regex:'k1_kstrongest|k3_surface|k5_register' --launch-skip 29 --launch-count 3 \
    -f -o gpurun_out/prof_r02t python bench.py --serial --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_r02t.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02t.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list_r02t.log 2>&1
ls -la gpurun_out/prof_r02t.ncu-rep; tail -2 gpurun_out/ncu_full_r02t.log; wc -l gpurun_out/launches_r02t.csv
