#!/bin/bash
O=gpurun_out/exp38; mkdir -p $O
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
SHAPE="4096 4096 1"; run PDWT_PATH=fused
SHAPE="4095 4097 1"; run PDWT_X=0
SHAPE="4094 4098 1"; run PDWT_X=0
SHAPE="4096 4100 1"; run PDWT_X=0
SHAPE="4095 4096 1"; run PDWT_X=0
