#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/k1_widths_r02l.txt
for l in k1w8 k1w8b; do
python - $l <<'PY' >> gpurun_out/k1_widths_r02l.txt 2>&1
import os, sys, runpy
sys.path.insert(0, os.getcwd())
from cfear_radarodometry_code_public_b200 import capi
capi.LIB_PATH = os.path.abspath("profiles/ab/libcfear_%s.so" % sys.argv[1])
print(sys.argv[1])
sys.argv = ["k1_widths.py", "256"]
runpy.run_path("profiles/k1_widths.py", run_name="__main__")
PY
done
cat gpurun_out/k1_widths_r02l.txt
bash profiles/gpu_tests.sh r02l
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_r02l_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|smoke ok" gpurun_out/san_r02l_$tool.txt | tail -3
done
python profiles/run_configs.py > gpurun_out/configs_r02l.jsonl 2> gpurun_out/configs_r02l.err; cut -c1-600 gpurun_out/configs_r02l.jsonl; tail -3 gpurun_out/configs_r02l.err
