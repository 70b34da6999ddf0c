#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
timeout 500 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fwd', d['roofline']['frac'], d['roofline']['whole_op']['frac'], 'bwd', d['roofline_bwd']['frac']); f=d['e2e_fresh_batches']; print('fresh', f['ms_per_step'], f['ms_per_step_same_workload_one_cached_batch'])"; tail -3 gpurun_out/r2y_bench.err
timeout 300 python -c "
import torch, bench
from fvta_memexqa_b200 import _lib
r = bench.attention_roofline_bwd(torch.device('cuda:0'), path=_lib.PATH_FFMA)
print('CUDA-core backward at C3-max: %.1f us frac %.3f' % (r['launch_ms']*1e3, r['frac']))" 2>&1 | tail -1
timeout 300 python tools/lstm_bench.py 2>&1 | tail -10
FVTA_LIB_PATH=$PWD/_ab/libfvta_trace.so timeout 300 python tools/bt_trace.py > gpurun_out/bt_trace.log 2>&1; tail -7 gpurun_out/bt_trace.log
