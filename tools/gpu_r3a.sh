#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "step_handle or pipeline" > gpurun_out/r3a_t.log 2>&1; grep -E "Error|error|passed|failed|assert" gpurun_out/r3a_t.log | head -10
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','final_loss')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_spread_rank0']); f=d['e2e_fresh_batches']; print('fresh', f['ms_per_step'], f['ms_per_step_same_workload_one_cached_batch'], f['final_loss'])"; tail -3 gpurun_out/r3a_bench.err
