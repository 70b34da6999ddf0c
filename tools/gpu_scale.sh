#!/bin/bash
# 1/2/4/8-GPU bench lines on one box: weak scaling (fused peer-memory collective, and NCCL for the A/B at 8),
# then the strong-scaling C4 point (global batch 256).
mkdir -p gpurun_out
run() { # N extra-args tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu-baseline "$@" > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $n --steps 200 --warmup 10 "$@" > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  fi
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_${tag}.json').read().strip().splitlines()[-1])
    print('  ', d['n_gpus'], 'value %.0f' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['scaling'], d['collective'][:40])
except Exception as e:
    print('  no line:', e)
PY
}
run 8 w8_fused
run 8 w8_nccl --nccl
run 4 w4_fused
run 2 w2_fused
run 1 w1
run 8 s8_g256 --global-batch 256
run 4 s4_g256 --global-batch 256
run 2 s2_g256 --global-batch 256
run 1 s1_g256 --global-batch 256
timeout 300 python -m pytest tests/test_model_gpu.py -q -k "two_gpu" > gpurun_out/scale_pytest.log 2>&1; tail -2 gpurun_out/scale_pytest.log
