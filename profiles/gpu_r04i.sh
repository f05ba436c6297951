#!/bin/bash
tag=r04i; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_$tag.txt; tail -3 gpurun_out/pytest_$tag.txt
python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc --serial > gpurun_out/bench_serial_$tag.json 2> gpurun_out/bench_serial_$tag.err; tail -2 gpurun_out/bench_serial_$tag.err
python -c "
import json
for f in ('gpurun_out/bench_$tag.json','gpurun_out/bench_serial_$tag.json'):
    d=json.load(open(f)); r=d['roofline']
    print('%s value %.0f (%.4f ms/step) roofline %s frac %.4f alone %s whole %.3f'%(d['config']['api'][:40], d['value'], d['ms_per_step'], r['kernel'], r['frac'], r['stage_ms_alone'], r['whole_path']['frac']))"
CFEAR_K5_FORM=0 python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc --serial --steps 100 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('serial, K5 forced to the 128-thread form: %.0f (%.4f ms/step) alone %s'%(d['value'], d['ms_per_step'], d['roofline']['stage_ms_alone']))"
