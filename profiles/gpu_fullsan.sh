#!/bin/bash
# The whole GPU suite, then the whole GPU suite under compute-sanitizer (memcheck, synccheck, initcheck) and the
# registration-heavy files under racecheck.  usage: gpu_fullsan.sh TAG
tag=$1; mkdir -p gpurun_out
bash profiles/gpu_tests.sh $tag
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -1; done
for tool in memcheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/san_${tag}_suite_$tool.txt 2>&1
  echo "suite $tool rc=$?"; grep -E "SUMMARY|passed|failed" gpurun_out/san_${tag}_suite_$tool.txt | tail -3
done
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -x -q -p no:cacheprovider > gpurun_out/san_${tag}_suite_racecheck.txt 2>&1
echo "configs+golden racecheck rc=$?"; grep -E "SUMMARY|passed|failed" gpurun_out/san_${tag}_suite_racecheck.txt | tail -3
