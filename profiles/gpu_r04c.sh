#!/bin/bash
# GPU suite with every K3 / K5 test in both launch forms, then compute-sanitizer on smoke() in both forms and
# synccheck / racecheck on the golden-fixture file (both forms).  usage: gpu_r04c.sh TAG
tag=$1; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_$tag.txt
tail -4 gpurun_out/pytest_$tag.txt
for w in 1 0; do
  for tool in memcheck racecheck synccheck initcheck; do
    CFEAR_K3_WIDE=$w CFEAR_K5_WIDE=$w timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tag}_wide${w}_$tool.txt 2>&1
    echo "smoke wide=$w $tool rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/san_${tag}_wide${w}_$tool.txt | tail -2
  done
done
for tool in synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -p no:cacheprovider > gpurun_out/san_${tag}_golden_$tool.txt 2>&1
  echo "golden $tool rc=$?"; grep -E "SUMMARY|passed|failed" gpurun_out/san_${tag}_golden_$tool.txt | tail -3
done
