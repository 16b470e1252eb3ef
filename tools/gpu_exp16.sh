#!/bin/bash
# non-separable inverse: filter products as kernel parameters (uniform constant loads instead of shared-memory broadcasts)
O=gpurun_out/exp16; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 200 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; cut -c1-200 $O/configs.jsonl | head -3
