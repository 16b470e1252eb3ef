#!/bin/bash
T="timeout 120 python tools/time_step.py"
{
$T
PDWT_TWOPERSM=0 $T
$T 2048 2048 1
PDWT_TWOPERSM=0 $T 2048 2048 1
} 2>&1 | grep -v "^$" | tee gpurun_out/exp9.txt
