#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --attn --steps 30 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print('tile ms', d['launch_ms'], 'min', d['launch_ms_min'], 'frac', d['frac'], 'whole op ms', d['whole_op']['launch_ms'], 'frac', d['whole_op']['frac'], d['clocks'], d['parity'])"
echo skip
