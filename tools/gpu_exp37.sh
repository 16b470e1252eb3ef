#!/bin/bash
O=gpurun_out/exp37; mkdir -p $O
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 1" "4096 4096 8"; do
  run PDWT_X=0
  run PDWT_INV_SMEM_KB=18
  run PDWT_INV_SMEM_KB=22
  run PDWT_INV_SMEM_KB=27
  run PDWT_INV_SMEM_KB=36
  run PDWT_TM=64
  run PDWT_TM=48
done
