#!/bin/bash
# Round-end evidence run (one GPU): parity tests, smoke, both bench arms, ncu launch list and one
# full capture of the dominant attention kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 450 --csv \
  --log-file gpurun_out/launches_${TAG}_bench_eager.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline \
  > gpurun_out/bench_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -f \
  -o gpurun_out/attn_fwd_tc_$TAG python tools/profile_step.py --attn --steps 2 > gpurun_out/attn_ncu_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
