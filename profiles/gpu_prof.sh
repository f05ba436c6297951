#!/bin/bash
# clock64 phase probes of K3 / K5 (a -DCFEAR_K5_PROFILE -DCFEAR_K3_PROFILE build in profiles/ab/libcfear_prof.so)
mkdir -p gpurun_out
python profiles/ab_stage.py --make /tmp/b.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_prof.so --batch /tmp/b.npz --prof --steps 10 > gpurun_out/prof_$1.txt 2>&1
grep -v "^K3 scan" gpurun_out/prof_$1.txt | tail; grep "^K3 scan" gpurun_out/prof_$1.txt | tail -4
