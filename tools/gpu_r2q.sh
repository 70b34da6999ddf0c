#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/r2q_bwd_launches.csv \
  python -c "
import torch, bench
print(bench.attention_roofline_bwd(torch.device('cuda:0'), iters=1))" > gpurun_out/r2q.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2q_bwd_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[hdr+1:hdr+25]:
    print(r[4][:60], r[-1])
PY
