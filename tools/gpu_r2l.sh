#!/bin/bash
# 2 GPUs: data-parallel parity (NCCL and fused peer-memory collective) + bench with both
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_model_gpu.py -q -k two_gpu -s > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2l_pytest.log
for mode in "" "--nccl"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 10 $mode > gpurun_out/r2l_bench2$mode.json 2> gpurun_out/r2l_bench2$mode.err; echo "bench2 $mode rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2l_bench2$mode.json')); print({k: d[k] for k in ('value','ms_per_step','collective')}, 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/r2l_bench2$mode.err
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2l_bench1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench1.json')); print('1 GPU', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
