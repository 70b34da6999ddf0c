#!/bin/bash
# round 2, call A: hardware probe + parity of the new tcgen05 forward
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 60 tools/umma_probe.bin > gpurun_out/r2a_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2a_probe.log
cat gpurun_out/r2a_probe.log
timeout 900 python -m pytest tests/test_attn_tc_gpu.py -q -s -x > gpurun_out/r2a_tc_tests.log 2>&1; echo "tc rc=$?"
tail -40 gpurun_out/r2a_tc_tests.log
