#!/bin/bash
# round-end validation of HEAD on one B200: GPU suite, smoke, default bench line, ncu (--set full of K1 / K3 / K5 + launch
# list), single-sequence / configs lines, sanitizers on smoke() with K5 forced to each launch form.  usage: gpu_final.sh TAG
tag=$1; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_$tag.txt; tail -3 gpurun_out/pytest_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); r=d['roofline']; e=d['e2e']
print('value %.0f (%.4f ms/step) e2e %.0f frac_h2d %.3f roofline %s frac %.4f alone %s whole %.3f cpu %.0f/%.0f'%(d['value'], d['ms_per_step'], e['value'], e['frac_of_h2d_only'], r['kernel'], r['frac'], r['stage_ms_alone'], r['whole_path']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value']))"
bash profiles/gpu_ncu.sh $tag > /dev/null 2>&1; ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/launches_$tag.csv
( timeout 600 python profiles/run_configs.py ) > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; grep -o '"ms_per_frame_gpu_median": [0-9.]*' gpurun_out/configs_$tag.jsonl
for f in 0 1 2; do
  for tool in memcheck racecheck synccheck; do
    CFEAR_K5_FORM=$f timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tag}_form${f}_$tool.txt 2>&1
    echo "smoke K5 form $f $tool rc=$? $(grep -E 'SUMMARY' gpurun_out/san_${tag}_form${f}_$tool.txt | tail -1)"
  done
done
