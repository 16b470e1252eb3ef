#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
T="timeout 120 python tools/time_step.py"
{
$T
$T 2048 2048 64
$T 4096 4096 8
$T 2048 2048 1
PDWT_TM=16 $T
} 2>&1 | grep -v "^$" | tee gpurun_out/exp7.txt
