#!/bin/bash
# Round-end evidence run (one GPU): parity tests, smoke, both bench arms, ncu launch list and full captures
# of the attention tile kernels (C3-max) and of the step's other hot kernels (C2).  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 450 --csv \
  --log-file gpurun_out/launches_${TAG}_bench_eager.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline \
  > gpurun_out/bench_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 6 -c 1 -f \
  -o gpurun_out/attn_fwd_tc_$TAG python tools/profile_step.py --attn --steps 2 > gpurun_out/attn_ncu_$TAG.log 2>&1; echo "ncu fwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 3 -c 1 -f \
  -o gpurun_out/attn_bwd_tc_$TAG python -c "
import torch, bench
print(bench.attention_roofline_bwd(torch.device('cuda:0'), iters=1)['launch_ms'])" > gpurun_out/attn_bwd_ncu_$TAG.log 2>&1; echo "ncu bwd rc=$?"
# the step's other hot kernels at C2: one launch each of the recurrences (forward, backward), the GEMM, the
# attention tile kernels and the char-CNN, after 3 eager warm-up steps
timeout 600 ncu --set full --clock-control none -k regex:'bilstm_rec|lstm_rec|gemm_mma_kernel|attn_bwd_tc|attn_fwd_tc|charcnn' \
  -s 60 -c 24 -f -o gpurun_out/step_kernels_$TAG python tools/profile_step.py --steps 1 --warmup 3 \
  > gpurun_out/step_ncu_$TAG.log 2>&1; echo "ncu step rc=$?"
ls -la gpurun_out | tail -14
