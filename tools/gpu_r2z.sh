#!/bin/bash
mkdir -p gpurun_out
export FVTA_LIB_PATH=$PWD/_ab/libfvta_b.so
timeout 600 python -m pytest tests/test_attn_tc_gpu.py tests/test_attn_gpu.py -q -x > gpurun_out/r2z_t.log 2>&1; grep -E "Error|error|passed|failed" gpurun_out/r2z_t.log | head -10
timeout 300 python -c "
import torch, bench
r = bench.attention_roofline_bwd(torch.device('cuda:0'))
print('B bwd whole op us %.1f frac %.3f' % (r['launch_ms']*1e3, r['frac']))" 2>&1 | tail -2
unset FVTA_LIB_PATH
timeout 300 python -c "
import torch, bench
r = bench.attention_roofline_bwd(torch.device('cuda:0'))
print('A bwd whole op us %.1f frac %.3f' % (r['launch_ms']*1e3, r['frac']))" 2>&1 | tail -2
FVTA_LIB_PATH=$PWD/_ab/libfvta_trace.so timeout 300 python tools/bt_trace.py > gpurun_out/bt_trace.log 2>&1; tail -8 gpurun_out/bt_trace.log
