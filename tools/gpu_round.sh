#!/bin/bash
# one gpurun call: parity tests, smoke, micro-benchmark, both bench arms, ncu launch list (+ optional full capture)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
[ -x tools/bin/ubench_fma ] && tools/bin/ubench_fma > gpurun_out/ubench_fma.txt 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/ubench_fma.txt; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
