#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_tests.sh r03a
rm -f /tmp/ab2_ref.npz
bash profiles/gpu_ab2.sh r03a profiles/ab/libcfear_cur.so profiles/ab/libcfear_k1p.so profiles/ab/libcfear_k1q.so profiles/ab/libcfear_k1p5.so profiles/ab/libcfear_k1q5.so profiles/ab/libcfear_cur.so
python profiles/k1_widths.py 128 > gpurun_out/k1_widths_r03a.txt 2>&1; cat gpurun_out/k1_widths_r03a.txt
