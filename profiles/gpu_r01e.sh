#!/bin/bash
# one GPU call: parity tests on the new default library, then the A/B stage timings of the K5 variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r01e_gpu.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r01e_pytest.txt
python profiles/ab_stage.py --make /tmp/b.npz
{
python profiles/ab_stage.py --lib profiles/ab/libcfear_v4.so --batch /tmp/b.npz --save /tmp/ref.npz
python profiles/ab_stage.py --lib cfear_radarodometry_code_public_b200/libcfear_b200.so --batch /tmp/b.npz --ref /tmp/ref.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_v5_192.so --batch /tmp/b.npz --ref /tmp/ref.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_v5_256.so --batch /tmp/b.npz --ref /tmp/ref.npz
python profiles/ab_stage.py --lib profiles/ab/libcfear_v5_prof.so --batch /tmp/b.npz --ref /tmp/ref.npz --prof --steps 20
python profiles/ab_stage.py --lib profiles/ab/libcfear_v4.so --batch /tmp/b.npz --ref /tmp/ref.npz
} > gpurun_out/r01e_ab.txt 2>&1
python bench.py --steps 200 --cpu-seconds 6 > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err
tail -3 gpurun_out/r01e_pytest.txt; cat gpurun_out/r01e_ab.txt; cut -c1-600 gpurun_out/r01e_bench.json
