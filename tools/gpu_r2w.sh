#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_tc_gpu.py -q -x -s -k tied_columns > gpurun_out/r2w_t.log 2>&1; grep -E "tied columns|Error|error|passed|failed|assert" gpurun_out/r2w_t.log | head -40
