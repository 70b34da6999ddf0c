#!/bin/bash
# A/B of two library builds on the C2 step: per-kernel time of the attention tile kernels (ncu launch list)
mkdir -p gpurun_out
for v in A B; do
  if [ $v = A ]; then unset FVTA_LIB_PATH; else export FVTA_LIB_PATH=$PWD/_ab/libfvta_b.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_fwd_tc|attn_bwd_tc' -s 6 -c 6 --csv \
     --log-file gpurun_out/ab2_$v.csv python tools/profile_step.py --steps 3 --warmup 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/ab2_$v.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
print('$v', [(r[4][11:22], r[-1]) for r in rows[h+1:]])
PY
  timeout 300 python bench.py --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v step ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
