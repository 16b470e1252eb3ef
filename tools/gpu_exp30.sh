#!/bin/bash
O=gpurun_out/exp30; mkdir -p $O
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 8" "2048 2048 64"; do
  run PDWT_MULTI=0
  run PDWT_MULTI=0 PDWT_LOWOCC=1
  run PDWT_MULTI=0 PDWT_LOWOCC=1 PDWT_TH=56
  run PDWT_MULTI=0 PDWT_TH=56
  run PDWT_MULTI=0 PDWT_TH=64
  run PDWT_MULTI=0 PDWT_LOWOCC=1 PDWT_TH=64
  run PDWT_MULTI=0 PDWT_LOWOCC=1 PDWT_TH=128
  run PDWT_MULTI=0 PDWT_LOWOCC=1 PDWT_TWOPERSM=0
done
