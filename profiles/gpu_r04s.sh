#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mirror.py -m gpu -x -q 2>&1 | tail -4 )
( timeout 600 python profiles/run_configs.py ) > gpurun_out/configs_r04s.jsonl 2> gpurun_out/configs_r04s.err; grep -o '"ms_per_frame_gpu_median": [0-9.]*\|"max_pos_diff_vs_oracle_replay_m": [0-9.e-]*' gpurun_out/configs_r04s.jsonl
