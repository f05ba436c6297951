#!/bin/bash
# wide K3 (1024 threads, one CTA per SM, batches of <= one scan per SM): GPU suite + sequence replays with and without it
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_r04b.txt
tail -4 gpurun_out/pytest_r04b.txt
: > gpurun_out/replay_r04b.jsonl
for n in 1 8 32 148; do
  for w in 1 0; do
    echo "nseq $n k3wide $w" >> gpurun_out/replay_r04b.jsonl
    CFEAR_K3_WIDE=$w timeout 300 python replay.py --nseq $n --steps 24 >> gpurun_out/replay_r04b.jsonl 2>> gpurun_out/replay_r04b.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/replay_r04b.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['sequences_per_gpu'], round(d['scans_per_s_device_resident']), round(d['scans_per_s_host_images']), d['max_pos_err_vs_oracle_replay_m'])
    else: print(l.strip())
PY
( time timeout 600 python profiles/run_configs.py ) > gpurun_out/configs_r04b.jsonl 2> gpurun_out/configs_r04b.err
tail -3 gpurun_out/configs_r04b.err; cut -c1-400 gpurun_out/configs_r04b.jsonl
