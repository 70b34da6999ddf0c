#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -c "
import torch, bench
r = bench.attention_roofline_bwd(torch.device('cuda:0'), iters=10)
print('bwd whole op ms', r['launch_ms'], 'frac', r['frac'])" > gpurun_out/r2t_a.log 2>&1; tail -12 gpurun_out/r2t_a.log
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_attn_tc_gpu.py -q -x -k backward > gpurun_out/r2t_b.log 2>&1; grep -E "Error|error|passed|failed" gpurun_out/r2t_b.log | head -10
