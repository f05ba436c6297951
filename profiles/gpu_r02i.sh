#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_tests.sh r02i
bash profiles/gpu_ab2.sh r02i profiles/ab/libcfear_bar1.so profiles/ab/libcfear_bar3.so cfear_radarodometry_code_public_b200/libcfear_b200.so
python profiles/k1_widths.py 128 > gpurun_out/k1_widths_r02i.txt 2>&1; cat gpurun_out/k1_widths_r02i.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_r02i_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/san_r02i_$tool.txt | tail -3
done
