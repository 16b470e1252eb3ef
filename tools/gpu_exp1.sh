#!/bin/bash
# experiment batch 1: parity, then timing variants (one process each), bench lines, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
T="timeout 120 python tools/time_step.py"
{
$T
PDWT_PDL=1 $T
PDWT_PDL=2 $T
PDWT_TH=32 $T
PDWT_TH=16 $T
PDWT_TM=8 $T
PDWT_TM=32 $T
PDWT_PATH=fused PDWT_FUSED_TILE=0 $T
PDWT_PATH=fused PDWT_FUSED_TILE=1 $T
PDWT_SMALL_PX=300000 $T
PDWT_SMALL_PX=1100000 $T
PDWT_SMALL_PX=1100000 PDWT_PDL=2 $T
PDWT_SMALL_PX=1100000 PDWT_PDL=1 $T
PDWT_SMALL_PX=1100000 PDWT_FUSED_TILE=0 $T
$T 2048 2048 64
PDWT_PDL=2 $T 2048 2048 64
$T 4096 4096 8
$T 2048 2048 1
PDWT_SMALL_PX=300000 $T 2048 2048 1
} 2>&1 | grep -v "^$" | tee gpurun_out/exp1.txt
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-1500 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 6 -o gpurun_out/ncu_c2 python tools/prof_fwdinv.py 3 > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 6 -o gpurun_out/ncu_b8 python tools/prof_batch.py > gpurun_out/ncu_b8.log 2>&1; tail -2 gpurun_out/ncu_b8.log
ls -la gpurun_out
