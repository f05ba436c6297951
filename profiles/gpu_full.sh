#!/bin/bash
# parity tests on the default library, then A/B stage timings against a reference build
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest.txt
tail -3 gpurun_out/pytest.txt
bash profiles/gpu_ab.sh "$@"
