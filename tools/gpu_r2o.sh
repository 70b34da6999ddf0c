#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_tc_gpu.py -q -s -x -k "backward" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "tc bwd|passed|failed|Error|error" gpurun_out/r2o_pytest.log | tail -40
