#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
T="timeout 120 python tools/time_step.py"
{
$T
PDWT_PDL=0 $T
$T 2048 2048 64
$T 4096 4096 8
$T 2048 2048 1
$T 1024 1024 1
$T 512 512 128
} 2>&1 | grep -v "^$" | tee gpurun_out/exp3.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 12 -c 6 --csv --log-file gpurun_out/launches_c2.csv python tools/prof_fwdinv.py 4 > gpurun_out/ncu_l.log 2>&1; grep '^"' gpurun_out/launches_c2.csv | awk -F'","' '{print $5, $NF}' | cut -c1-80 | tail -6
