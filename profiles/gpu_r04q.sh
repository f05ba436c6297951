#!/bin/bash
# ncu --set full of the wide forms (one CTA per SM): stream-ordered 128-scan steps
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k1_kstrongest|k3_surface|k5_register' --launch-skip 29 --launch-count 3 \
    -f -o gpurun_out/prof_r04q python bench.py --nprob 128 --serial --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_r04q.log 2>&1
ls -la gpurun_out/prof_r04q.ncu-rep; tail -2 gpurun_out/ncu_full_r04q.log
