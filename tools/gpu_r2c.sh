#!/bin/bash
mkdir -p gpurun_out
FVTA_LIB_PATH=fvta_memexqa_b200/lib/libfvta_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/r2c_trace.log 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/r2c_trace.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 6 -c 1 -f \
  -o gpurun_out/attn_fwd_tc_r2c python tools/profile_step.py --attn --steps 4 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_pytest.log | tail -15
