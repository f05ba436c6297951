#!/bin/bash
# one ncu --set full capture of K1 / K3 / K5 in steady state (stream-ordered steps: ncu serialises the kernels anyway) +
# the launch list of a short default bench run.  usage: gpu_ncu.sh TAG
mkdir -p gpurun_out
# matching launches before the capture: 8 (keyframe sets: 4 x K1, K3) + 3 warm-up steps x 3 + 4 estimate steps x 3 = 29
ncu --set full --clock-control none --import-source on -k regex:'k1_kstrongest|k3_surface|k5_register' --launch-skip 29 --launch-count 3 \
    -f -o gpurun_out/prof_$1 python bench.py --serial --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list_$1.log 2>&1
ls -la gpurun_out/ | tail -8; tail -3 gpurun_out/ncu_full_$1.log
