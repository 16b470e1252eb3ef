#!/bin/bash
O=gpurun_out/exp46; mkdir -p $O
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | cut -c1-120 | tee -a $O/times.txt; }
export PDWT_INV_TMA=1
SHAPE="4096 4096 1"; run PDWT_TM=64; run PDWT_TM=128; run PDWT_TM=16
SHAPE="1024 1024 8"; run PDWT_X=0; run PDWT_TM=32
SHAPE="4096 4096 2"; run PDWT_X=0; run PDWT_TM=32
SHAPE="2048 2048 4"; run PDWT_TM=32
SHAPE="2048 2048 16"; run PDWT_TM=32
