#!/bin/bash
# quick verification on one GPU: the whole GPU suite and the bench line (no CPU baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/check_t.log 2>&1; grep -E "Error|error|passed|failed|assert" gpurun_out/check_t.log | head -10
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/check_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','final_loss')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step']); f=d['e2e_fresh_batches']; print('fresh', f.get('ms_per_step'), f.get('ms_per_step_same_workload_one_cached_batch'), f.get('error'))"; tail -3 gpurun_out/check_bench.err
