#!/bin/bash
# usage: gpu_ab5.sh TAG "LIBS" "INFLIGHTS"
tag=$1; mkdir -p gpurun_out; : > gpurun_out/ab5_$tag.txt
for lib in $2; do for n in $3; do
  python bench.py --lib profiles/ab/libcfear_$lib.so --inflight $n --batch-cache /tmp/bc --no-cpu --no-e2e --steps 200 > /tmp/o.json 2> /tmp/o.err || tail -3 /tmp/o.err >> gpurun_out/ab5_$tag.txt
  python -c "
import json; d=json.load(open('/tmp/o.json')); r=d['roofline']; print('$lib in flight $n: %.4f ms/step %.0f scans/s | alone'%(d['ms_per_step'], d['value']), {k[:2]: round(v,4) for k,v in r['stage_ms_alone'].items()}, 'in flight', {k[:2]: round(v,4) for k,v in r['stage_ms_in_flight'].items()})" >> gpurun_out/ab5_$tag.txt
done; done
cat gpurun_out/ab5_$tag.txt
