#!/bin/bash
# r02d: K3 variants (2 CTAs/SM layouts), 3 steps in flight, K5/K3 clock64 profile
bash profiles/gpu_ab2.sh r02d profiles/ab/libcfear_v9a.so profiles/ab/libcfear_k3n384.so profiles/ab/libcfear_k3n768.so profiles/ab/libcfear_k3n512.so profiles/ab/libcfear_k3n256.so
python bench.py --lib profiles/ab/libcfear_p3.so --inflight 3 --batch-cache /tmp/bc --no-cpu --no-e2e --steps 200 > gpurun_out/r02d_p3.json 2> gpurun_out/r02d_p3.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_p3.json')); print('3 in flight: %.4f ms/step %.0f scans/s'%(d['ms_per_step'], d['value']), d['roofline']['stage_ms_per_step'])" >> gpurun_out/ab2_r02d.txt
python profiles/ab_stage.py --make /tmp/b.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_prof.so --batch /tmp/b.npz --prof --steps 10 > gpurun_out/r02d_prof.txt 2>&1
tail -5 gpurun_out/ab2_r02d.txt; grep -v "^K3 scan" gpurun_out/r02d_prof.txt | tail; grep "^K3 scan" gpurun_out/r02d_prof.txt | tail -3
