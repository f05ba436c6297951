#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/san_$tool.txt | tail -3
done
