#!/bin/bash
mkdir -p gpurun_out
python profiles/ab_stage.py --make /tmp/b.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_v5_prof.so --batch /tmp/b.npz --prof --steps 20 > gpurun_out/prof.txt 2>&1
cat gpurun_out/prof.txt
