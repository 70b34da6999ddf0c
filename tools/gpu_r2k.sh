#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --attn --steps 30 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print('tile ms', d['launch_ms'], 'min', d['launch_ms_min'], 'frac', d['frac'], 'whole op ms', d['whole_op']['launch_ms'], 'frac', d['whole_op']['frac'], d['clocks'], d['parity'])"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_pytest.log
timeout 400 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_op']['frac'])"; tail -3 gpurun_out/r2k_bench.err
