#!/bin/bash
# r02e: steps in flight x K3 layout
mkdir -p gpurun_out; : > gpurun_out/ab_r02e.txt
for lib in k3n384 k3n768 k3n512; do for n in 2 3 4 6; do
  python bench.py --lib profiles/ab/libcfear_$lib.so --inflight $n --batch-cache /tmp/bc --no-cpu --no-e2e --steps 200 > /tmp/o.json 2> /tmp/o.err || tail -3 /tmp/o.err >> gpurun_out/ab_r02e.txt
  python -c "
import json; d=json.load(open('/tmp/o.json')); print('$lib in flight $n: %.4f ms/step %.0f scans/s'%(d['ms_per_step'], d['value']), {k: round(v,4) for k,v in d['roofline']['stage_ms_per_step'].items()})" >> gpurun_out/ab_r02e.txt
done; done
cat gpurun_out/ab_r02e.txt
