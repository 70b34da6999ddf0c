#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bilstm_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2m_pytest.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_op']['frac'])"; tail -3 gpurun_out/r2m_bench.err
timeout 300 python tools/lstm_bench.py 2>&1 | tail -8
