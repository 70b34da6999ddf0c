#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 3 -c 1 -f -o gpurun_out/attn_bwd_tc_r2r \
  python -c "
import torch, bench
print(bench.attention_roofline_bwd(torch.device('cuda:0'), iters=1))" > gpurun_out/r2r.log 2>&1; echo "rc=$?"
