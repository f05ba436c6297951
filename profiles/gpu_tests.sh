#!/bin/bash
# GPU parity suite; output tail into gpurun_out/pytest_$1.txt
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_$1.txt
tail -6 gpurun_out/pytest_$1.txt
