#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 3 -c 1 -o gpurun_out/r2_attn_bwd_tc -f \
  python -c "
import torch, bench
print(bench.attention_roofline_bwd(torch.device('cuda:0'), iters=1)['launch_ms'])" > gpurun_out/r2v.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2v.log
