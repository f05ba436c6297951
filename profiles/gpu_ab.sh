#!/bin/bash
# usage: gpu_ab.sh REF_LIB [LIB...]   -- a *_prof.so library is run with --prof
mkdir -p gpurun_out
python profiles/ab_stage.py --make /tmp/b.npz
ref=$1; shift
{
python profiles/ab_stage.py --lib $ref --batch /tmp/b.npz --save /tmp/ref.npz
for l in "$@"; do
  case $l in
    *_prof*.so) python profiles/ab_stage.py --lib $l --batch /tmp/b.npz --ref /tmp/ref.npz --prof --steps 20 ;;
    *) python profiles/ab_stage.py --lib $l --batch /tmp/b.npz --ref /tmp/ref.npz ;;
  esac
done
} > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
