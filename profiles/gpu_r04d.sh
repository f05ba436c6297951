#!/bin/bash
# final single-GPU measurement of the round: default bench line, ncu --set full of K1 / K3 / K5, launch list.  usage: gpu_r04d.sh TAG
tag=$1; mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); r=d['roofline']; e=d['e2e']
print('value %.0f (%.4f ms/step) e2e %.0f frac_h2d %.3f roofline %s frac %.4f alone %s whole %.3f cpu %.0f/%.0f'%(d['value'], d['ms_per_step'], e['value'], e['frac_of_h2d_only'], r['kernel'], r['frac'], r['stage_ms_alone'], r['whole_path']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value']))"
bash profiles/gpu_ncu.sh $tag
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$tag.json
