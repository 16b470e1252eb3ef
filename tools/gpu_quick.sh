#!/bin/bash
# quick GPU iteration: parity tests, then our bench arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
