#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_fwd_tc -s 4 -c 1 -f \
  -o gpurun_out/lstm_tc_r2n python tools/lstm_bench.py > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2n_ncu.log
