#!/bin/bash
mkdir -p gpurun_out
bash profiles/gpu_ab.sh profiles/ab/libcfear_bar1.so profiles/ab/libcfear_bar0.so profiles/ab/libcfear_bar2.so
for b in 0 2; do
  CFEAR_LIB=profiles/ab/libcfear_bar$b.so timeout 600 compute-sanitizer --tool synccheck python - > gpurun_out/sync_bar$b.txt 2>&1 <<PY
import os, numpy as np
from cfear_radarodometry_code_public_b200 import capi, workload
capi.LIB_PATH = os.path.abspath(os.environ["CFEAR_LIB"])
nprob, K = 2, 2
b = workload.make_batch(nprob, K, seed0=100, workers=1)
ctx = capi.Context(device=0, max_batch=nprob, max_cellsets=nprob * (K + 1), max_keyframes=K, **workload.CFEAR3)
kf = np.arange(nprob * K, dtype=np.int32).reshape(nprob, K); cur = (nprob * K + np.arange(nprob)).astype(np.int32)
for i in range(K): ctx.scans_to_cells_batch(b["kf_polar"][:, i], None, kf[:, i])
out = ctx.odometry_step_batch(b["polar"], b["mot"], kf, cur, b["poses"])
print("ran", out["stats"]["outer_iterations"])
PY
  echo "bar$b:"; grep -E "SUMMARY|ran" gpurun_out/sync_bar$b.txt | tail -2
done
