#!/bin/bash
# usage: gpu_env_ab.sh TAG VAR "VALUE1" "VALUE2" ...  -- bench.py with the shipped library under VAR=VALUE, one summary line each
tag=$1; var=$2; shift; shift
mkdir -p gpurun_out
: > gpurun_out/envab_$tag.txt
for v in "$@"; do
  env $var="$v" python bench.py --batch-cache /tmp/bc --no-cpu --no-e2e --steps 200 $EXTRA > /tmp/envab.json 2> /tmp/envab.err || { echo "$v FAILED"; tail -3 /tmp/envab.err; } >> gpurun_out/envab_$tag.txt
  python - "$var=$v" >> gpurun_out/envab_$tag.txt <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/envab.json")); r = d["roofline"]
    print("%-26s %.4f ms/step (%.0f scans/s) | alone K1 %.4f K3 %.4f K5 %.4f | in-flight K1 %.4f K3 %.4f K5 %.4f" % (
        sys.argv[1], d["ms_per_step"], d["value"], *[r["stage_ms_alone"][k] for k in ("k1_kstrongest", "k3_surface_points", "k5_register")],
        *[r["stage_ms_in_flight"][k] for k in ("k1_kstrongest", "k3_surface_points", "k5_register")]))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
cat gpurun_out/envab_$tag.txt
