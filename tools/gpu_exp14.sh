#!/bin/bash
# non-separable inverse at 3 CTAs per SM (80 registers)
O=gpurun_out/exp14; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 200 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; cut -c1-200 $O/configs.jsonl | head -3
PDWT_NS_RG=2 timeout 100 python tools/bench_configs.py 2>/dev/null | head -3 | cut -c1-200
