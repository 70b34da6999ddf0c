#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --attn --steps 30 > gpurun_out/r2g_attn.log 2>&1; echo "attn rc=$?"; tail -2 gpurun_out/r2g_attn.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 6 -c 1 -f \
  -o gpurun_out/attn_fwd_tc_r2g python tools/profile_step.py --attn --steps 2 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu full rc=$?"
