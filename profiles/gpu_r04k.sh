#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r04k.txt
for f in 4 5 6 8; do
python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc --inflight $f 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('inflight $f: %.0f (%.4f ms/step) steps_in_flight %s'%(d['value'], d['ms_per_step'], d['config']['steps_in_flight']))" >> gpurun_out/r04k.txt
done
cat gpurun_out/r04k.txt
