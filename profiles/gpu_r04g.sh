#!/bin/bash
# small batches through the overlapped submit path (five steps in flight): wide vs batch forms of K3 / K5
mkdir -p gpurun_out; : > gpurun_out/r04g.txt
for n in 32 64 128 148; do
  for w in 1 0; do
    CFEAR_K3_WIDE=$w CFEAR_K5_WIDE=$w python bench.py --nprob $n --batch-cache /tmp/bc$n --no-cpu --no-e2e --steps 200 > /tmp/g.json 2>/tmp/g.err || tail -3 /tmp/g.err
    python -c "
import json; d=json.load(open('/tmp/g.json')); r=d['roofline']
print('nprob $n wide $w: overlapped %.4f ms/step %.0f scans/s | alone K1 %.4f K3 %.4f K5 %.4f'%(d['ms_per_step'], d['value'], *[r['stage_ms_alone'][k] for k in ('k1_kstrongest','k3_surface_points','k5_register')]))" >> gpurun_out/r04g.txt
  done
done
cat gpurun_out/r04g.txt
