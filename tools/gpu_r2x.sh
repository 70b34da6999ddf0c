#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -k "pipeline" > gpurun_out/r2x_t.log 2>&1; grep -E "Error|error|passed|failed|assert" gpurun_out/r2x_t.log | head -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']); print(d.get('e2e_fresh_batches'))"; tail -5 gpurun_out/r2x_bench.err
