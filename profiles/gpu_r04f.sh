#!/bin/bash
# 2-GPU check of the round's final build: bench.py (both arms) and replay.py under torch.distributed.run
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29520 bench.py --gpus 2 --steps 200 --warmup 3 > gpurun_out/bench_2gpu_r04f.json 2> gpurun_out/bench_2gpu_r04f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_r04f.json')); e=d['e2e']
print('N=2 value %.0f (%.4f ms/step) e2e %.0f (%.3f ms/step) h2d_only %.3f ms/step %.1f GB/s/GPU frac %.3f clocks %s'%(d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['h2d_only']['ms_per_step'], e['h2d_only']['gbytes_per_s_per_gpu'], e['frac_of_h2d_only'], d['clocks']))" || tail -5 gpurun_out/bench_2gpu_r04f.err
$TR --nproc-per-node 2 --master-port 29521 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-160
$TR --nproc-per-node 2 --master-port 29530 replay.py --nseq 32 --steps 24 --check 1 2> gpurun_out/replay_2gpu_r04f.err | grep scans_per_s > gpurun_out/replay_2gpu_r04f.jsonl
python -c "
import json
for l in open('gpurun_out/replay_2gpu_r04f.jsonl'):
    d=json.loads(l); print('replay N=%d resident %.0f host %.0f scans/s parity %.2e'%(d['n_gpus'], d['scans_per_s_device_resident'], d['scans_per_s_host_images'], d['max_pos_err_vs_oracle_replay_m']))"
