#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_ab2.sh r02k profiles/ab/libcfear_bar1.so profiles/ab/libcfear_loop4.so profiles/ab/libcfear_bar3.so
for l in loop4 k1u5; do
python - $l <<'PY' >> gpurun_out/k1_widths_r02k.txt 2>&1
import os, sys, runpy
sys.path.insert(0, os.getcwd())
from cfear_radarodometry_code_public_b200 import capi
capi.LIB_PATH = os.path.abspath("profiles/ab/libcfear_%s.so" % sys.argv[1])
print(sys.argv[1])
sys.argv = ["k1_widths.py", "256"]
runpy.run_path("profiles/k1_widths.py", run_name="__main__")
PY
done
cat gpurun_out/k1_widths_r02k.txt
