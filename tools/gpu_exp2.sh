#!/bin/bash
# experiment batch 2: parity, then timing of the new defaults against forced variants
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
T="timeout 120 python tools/time_step.py"
{
$T
PDWT_PDL=0 $T
PDWT_TH=64 PDWT_TM=16 $T
PDWT_TH=28 $T
PDWT_TH=116 $T
PDWT_TM=16 $T
PDWT_SMALL_PX=300000 $T
PDWT_SMALL_PX=1100000 $T
$T 2048 2048 64
PDWT_TH=64 $T 2048 2048 64
$T 4096 4096 8
$T 2048 2048 1
$T 1024 1024 1
$T 512 512 128
} 2>&1 | grep -v "^$" | tee gpurun_out/exp2.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-400 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 12 -c 12 --csv --log-file gpurun_out/launches_c2.csv python tools/prof_fwdinv.py 4 > gpurun_out/ncu_l.log 2>&1; grep -v "^==" gpurun_out/launches_c2.csv | cut -d, -f5,12- | cut -c1-120 | tail -13
