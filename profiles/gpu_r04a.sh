#!/bin/bash
# wide K5 (384 threads, one CTA per SM, batches of <= one problem per SM): GPU suite + sequence replays with and without it
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_r04a.txt
tail -4 gpurun_out/pytest_r04a.txt
: > gpurun_out/replay_r04a.jsonl
for n in 1 8 32 148; do
  for w in 1 0; do
    echo "nseq $n wide $w" >> gpurun_out/replay_r04a.jsonl
    CFEAR_K5_WIDE=$w timeout 300 python replay.py --nseq $n --steps 24 >> gpurun_out/replay_r04a.jsonl 2>> gpurun_out/replay_r04a.err
  done
done
cut -c1-20,160-330 gpurun_out/replay_r04a.jsonl
