#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest.log
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'fwd', d['roofline']['frac'], d['roofline']['whole_op']['frac'], 'bwd', d['roofline_bwd']['frac'], d['roofline_bwd']['launch_ms'])"; tail -3 gpurun_out/r2s_bench.err
