#!/bin/bash
# SWT kernels: packed exact adds, interleaved pair tiles, 4 outputs per thread in the row passes
O=gpurun_out/exp11; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 200 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; cut -c1-200 $O/configs.jsonl | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_swt" -s 8 -c 8 -o $O/ncu_swt python tools/prof_c3c4.py > $O/ncu_swt.log 2>&1
ncu -i $O/ncu_swt.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_swt_summary.txt; cat $O/ncu_swt_summary.txt
