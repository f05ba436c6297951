#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r04n.txt
b() { python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$1: %.0f (%.4f ms/step) in-flight %s'%(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['stage_ms_in_flight'].items()}))" >> gpurun_out/r04n.txt; }
b first
b second
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
b after_parity_tests
timeout 600 python -m pytest tests/test_gpu_mirror.py -m gpu -x -q 2>&1 | tail -1
b after_mirror_tests
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -1
b after_configs_tests
sleep 15
b after_sleep
nvidia-smi --query-gpu=temperature.gpu,temperature.memory,power.draw,clocks.mem,clocks.sm --format=csv >> gpurun_out/r04n.txt
cat gpurun_out/r04n.txt
