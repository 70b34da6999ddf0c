#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q -k "two_gpu" > gpurun_out/scale_pytest.log 2>&1; tail -2 gpurun_out/scale_pytest.log
for tag in fused nccl; do
  extra=""; [ $tag = nccl ] && extra="--nccl"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus 2 --steps 300 --warmup 10 $extra > gpurun_out/scale2_$tag.json 2> gpurun_out/scale2_$tag.err
  python -c "
import json; d=json.loads(open('gpurun_out/scale2_$tag.json').read().strip().splitlines()[-1]); print('$tag 2 value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
timeout 300 python bench.py --gpus 1 --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('w1 value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
