#!/bin/bash
# 8-GPU check of the data-parallel step: fused peer-memory collective vs NCCL, plus the 2-GPU parity tests
mkdir -p gpurun_out
run() { local n=$1; shift; local tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $n --steps 200 --warmup 10 "$@" > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_${tag}.json').read().strip().splitlines()[-1])
    print('  ', d['n_gpus'], 'value %.0f' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['collective'][:40])
except Exception as e:
    print('  no line:', e)
PY
}
timeout 300 python -m pytest tests/test_model_gpu.py -q -k "two_gpu" > gpurun_out/scale_pytest.log 2>&1; tail -2 gpurun_out/scale_pytest.log
run 8 w8_fused
run 8 w8_nccl --nccl
run 4 w4_fused
run 2 w2_fused
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/scale_w1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/scale_w1.json').read().strip().splitlines()[-1]); print('w1 value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
