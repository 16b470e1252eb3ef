#!/bin/bash
mkdir -p gpurun_out
T="timeout 120 python tools/time_step.py"
{
$T
$T 2048 2048 1
} 2>&1 | grep -v "^$" | tee gpurun_out/exp4.txt
timeout 600 python tools/bench_configs.py 2>&1 | grep -v "^$" | tee gpurun_out/configs.txt
