#!/bin/bash
mkdir -p gpurun_out
echo "softmax top (default)"; timeout 300 python tools/profile_step.py --attn --steps 30 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(d['launch_ms'], d['launch_ms_min'], d['frac'], d['whole_op']['launch_ms'], d['clocks'])"
echo "split top"; FVTA_LIB_PATH=fvta_memexqa_b200/lib/libfvta_splittop.so timeout 300 python tools/profile_step.py --attn --steps 30 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(d['launch_ms'], d['launch_ms_min'], d['frac'], d['whole_op']['launch_ms'], d['clocks'])"
