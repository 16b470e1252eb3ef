#!/bin/bash
O=gpurun_out/exp28; mkdir -p $O
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 1" "4096 4096 8" "2048 2048 64"; do
  run PDWT_MULTI=1
  run PDWT_TH=64 PDWT_TM=32
done
