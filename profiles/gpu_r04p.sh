#!/bin/bash
# 8-GPU line of the round's final build (bench.py, both device-resident and e2e arms)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 200 --warmup 3 > gpurun_out/bench_8gpu_r04p.json 2> gpurun_out/bench_8gpu_r04p.err
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_r04p.json')); e=d['e2e']
print('N=8 value %.0f (%.4f ms/step) e2e %.0f (%.3f ms/step) h2d_only %.3f ms/step %.1f GB/s/GPU frac %.3f clocks %s'%(d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['h2d_only']['ms_per_step'], e['h2d_only']['gbytes_per_s_per_gpu'], e['frac_of_h2d_only'], d['clocks']))" || tail -5 gpurun_out/bench_8gpu_r04p.err
