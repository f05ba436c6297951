#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r04j.txt
for i in 1 2 3; do
python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('default: %.0f (%.4f ms/step) clocks %s alone %s'%(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['stage_ms_alone']))" >> gpurun_out/r04j.txt
done
CFEAR_K5_FORM=0 python bench.py --no-cpu --no-e2e --batch-cache /tmp/bc 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('form0: %.0f (%.4f ms/step) clocks %s alone %s'%(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['stage_ms_alone']))" >> gpurun_out/r04j.txt
cat gpurun_out/r04j.txt
