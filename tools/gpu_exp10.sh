#!/bin/bash
# flat persistent element-wise kernels + host publication of the norms: parity, sequence timing, ncu of the three kernels
O=gpurun_out/exp10; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 120 python tools/prof_seq.py > $O/sequence.txt 2>&1; cat $O/sequence.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_reduce|k_threshold" -c 6 -o $O/ncu_elem python tools/prof_elem.py > $O/ncu_elem.log 2>&1
ncu -i $O/ncu_elem.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_elem_summary.txt; cat $O/ncu_elem_summary.txt
timeout 200 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; cut -c1-260 $O/configs.jsonl
