#!/bin/bash
# round 2, call B: roofline of the new forward, whole GPU suite, bench, ncu capture
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --attn --steps 10 > gpurun_out/r2b_attn.log 2>&1; echo "attn rc=$?"; tail -5 gpurun_out/r2b_attn.log
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
timeout 400 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cat gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 3 -c 1 -f \
  -o gpurun_out/attn_fwd_tc_r2b python tools/profile_step.py --attn --steps 2 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu full rc=$?"
