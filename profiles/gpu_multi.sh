#!/bin/bash
# one 8-GPU box: concurrent H2D ceiling, bench.py at 4 / 8 ranks, replay.py (configs[4]) at 2 / 4 / 8 ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/multi_topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> gpurun_out/multi_topo.txt
: > gpurun_out/multi_h2d.jsonl
python profiles/h2d_bench.py >> gpurun_out/multi_h2d.jsonl 2>/dev/null
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 29511 profiles/h2d_bench.py 2>/dev/null | grep gbs_ >> gpurun_out/multi_h2d.jsonl
done
$TR --nproc-per-node 8 --master-port 29512 profiles/h2d_bench.py --streams 2 2>/dev/null | grep gbs_ >> gpurun_out/multi_h2d.jsonl
$TR --nproc-per-node 8 --master-port 29513 profiles/h2d_bench.py --numa 2>/dev/null | grep gbs_ >> gpurun_out/multi_h2d.jsonl
cat gpurun_out/multi_h2d.jsonl
for n in 4 8; do
  $TR --nproc-per-node $n --master-port 29520 bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/multi_bench_$n.json 2> gpurun_out/multi_bench_$n.err
  python -c "
import json; d=json.load(open('gpurun_out/multi_bench_$n.json')); e=d['e2e']
print('N=$n value %.0f (%.4f ms/step) e2e %.0f (%.3f ms/step) h2d_only %.3f ms/step %.1f GB/s/GPU frac %.3f clocks %s'%(d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['h2d_only']['ms_per_step'], e['h2d_only']['gbytes_per_s_per_gpu'], e['frac_of_h2d_only'], d['clocks']))" || tail -3 gpurun_out/multi_bench_$n.err
done
: > gpurun_out/multi_replay.jsonl
python replay.py --nseq 32 --steps 24 --check 1 >> gpurun_out/multi_replay.jsonl 2> gpurun_out/multi_replay.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 29530 replay.py --nseq 32 --steps 24 --check 1 2>> gpurun_out/multi_replay.err | grep scans_per_s >> gpurun_out/multi_replay.jsonl
done
python -c "
import json
for l in open('gpurun_out/multi_replay.jsonl'):
    d=json.loads(l); print('replay N=%d resident %.0f host %.0f scans/s parity %.2e'%(d['n_gpus'], d['scans_per_s_device_resident'], d['scans_per_s_host_images'], d['max_pos_err_vs_oracle_replay_m']))"
