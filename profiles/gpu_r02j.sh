#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_ab2.sh r02j profiles/ab/libcfear_bar1.so profiles/ab/libcfear_loopk1.so
cp profiles/ab/libcfear_loopk1.so /tmp/keep.so
python - <<'PY' > gpurun_out/k1_widths_r02j.txt 2>&1
import os, sys, runpy
sys.path.insert(0, os.getcwd())
from cfear_radarodometry_code_public_b200 import capi
capi.LIB_PATH = os.path.abspath("profiles/ab/libcfear_loopk1.so")
sys.argv = ["k1_widths.py", "128"]
runpy.run_path("profiles/k1_widths.py", run_name="__main__")
PY
cat gpurun_out/k1_widths_r02j.txt
