#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
T="timeout 120 python tools/time_step.py"
{
$T
PDWT_LOWOCC=0 $T
$T 2048 2048 1
PDWT_LOWOCC=0 $T 2048 2048 1
$T 4096 4096 8
} 2>&1 | grep -v "^$" | tee gpurun_out/exp8.txt
