#!/bin/bash
# tests, bench line, ncu full capture + launch list for tag $1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest.txt
tail -3 gpurun_out/pytest.txt
python bench.py --cpu-seconds 8 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json"))
print("value %.0f scans/s  ms/step %.4f  e2e %.0f (%.3f ms/step)  cpu %.0f (%d cores)  roofline %s frac %.4f  stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["stage_ms_per_step"]))
PY
bash profiles/gpu_ncu.sh $1 > /dev/null 2>&1
ls -la gpurun_out | tail -6
