#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --attn --steps 10 > gpurun_out/r2f_attn.log 2>&1; echo "attn rc=$?"; tail -2 gpurun_out/r2f_attn.log | cut -c1-400
FVTA_LIB_PATH=fvta_memexqa_b200/lib/libfvta_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/r2f_trace.log 2>&1; echo "trace rc=$?"; tail -22 gpurun_out/r2f_trace.log
timeout 600 python -m pytest tests/test_attn_tc_gpu.py -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
