#!/bin/bash
# A/B/C of library builds on the forward/backward attention roofline (same box, interleaved)
mkdir -p gpurun_out
for rep in 1 2; do
for v in a b c; do
  export FVTA_LIB_PATH=$PWD/_ab/libfvta_$v.so
  timeout 300 python -c "
import torch, bench
r = bench.attention_roofline(torch.device('cuda:0'))
b = bench.attention_roofline_bwd(torch.device('cuda:0'))
print('$v fwd tile us %.1f frac %.3f whole %.1f | bwd whole us %.1f frac %.3f' % (r['launch_ms']*1e3, r['frac'], r['whole_op']['launch_ms']*1e3, b['launch_ms']*1e3, b['frac']))" 2>&1 | tail -1
done; done
