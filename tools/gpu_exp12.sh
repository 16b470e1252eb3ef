#!/bin/bash
# non-separable inverse: 128-bit window loads, one filter fetch per two positions
O=gpurun_out/exp12; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 200 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; cut -c1-200 $O/configs.jsonl | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_nonsep" -s 4 -c 4 -o $O/ncu_ns python tools/prof_c3c4.py > $O/ncu_ns.log 2>&1
ncu -i $O/ncu_ns.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_ns_summary.txt; cat $O/ncu_ns_summary.txt
