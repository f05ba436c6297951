#!/bin/bash
# usage: gpu_ab2.sh TAG LIB...   -- bench.py (overlapped value + per-kernel times alone) for each library, one summary line each
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/ab2_$tag.txt
for l in "$@"; do
  ref=""; [ -f /tmp/ab2_ref.npz ] && ref="--ref-poses /tmp/ab2_ref.npz" || ref="--save-poses /tmp/ab2_ref.npz"
  python bench.py --lib $l --batch-cache /tmp/bc --no-cpu --no-e2e --steps 200 $ref > /tmp/ab2.json 2> /tmp/ab2.err || { echo "$l FAILED"; tail -3 /tmp/ab2.err; } >> gpurun_out/ab2_$tag.txt
  python - "$l" >> gpurun_out/ab2_$tag.txt <<'PY'
import json, sys, hashlib
try:
    d = json.load(open("/tmp/ab2.json")); r = d["roofline"]; w = d["workload_stats"]
    print("%-34s overlapped %.4f ms/step (%.0f scans/s) | alone K1 %.4f K3 %.4f K5 %.4f | in-flight K1 %.4f K3 %.4f K5 %.4f | outer %.3f inner %.3f res %.1f poserr %.3e" % (
        sys.argv[1].split("/")[-1], d["ms_per_step"], d["value"], *[r["stage_ms_alone"][k] for k in ("k1_kstrongest", "k3_surface_points", "k5_register")],
        *[r["stage_ms_in_flight"][k] for k in ("k1_kstrongest", "k3_surface_points", "k5_register")],
        w["outer_iterations_mean"], w["inner_iterations_mean"], w["residuals_mean"], w["median_pos_err_vs_truth_m"]), d.get("ab_vs_ref", ""))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
cat gpurun_out/ab2_$tag.txt
