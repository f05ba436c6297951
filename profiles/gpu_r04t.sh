#!/bin/bash
bash profiles/gpu_ab.sh r04t base fastsc base fastsc
for l in base fastsc; do
  for n in 1 32; do
    CFEAR_B200_LIB=$PWD/profiles/ab/libcfear_$l.so timeout 300 python replay.py --nseq $n --steps 24 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$l nseq $n: %.0f scans/s resident, parity %.2e' % (d['scans_per_s_device_resident'], d['max_pos_err_vs_oracle_replay_m']))" | tee -a gpurun_out/ab2_r04t.txt
  done
done
