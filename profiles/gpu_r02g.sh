#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_tests.sh r02g
python replay.py --nseq 32 --steps 40 > gpurun_out/replay32_r02g.json 2> gpurun_out/replay_r02g.err; tail -2 gpurun_out/replay_r02g.err; cat gpurun_out/replay32_r02g.json
python replay.py --nseq 148 --steps 24 --check 2 > gpurun_out/replay148_r02g.json 2>> gpurun_out/replay_r02g.err; cat gpurun_out/replay148_r02g.json
