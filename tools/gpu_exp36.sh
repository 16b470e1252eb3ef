#!/bin/bash
O=gpurun_out/exp36; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -8 $O/pytest.log
python tools/bench_configs.py nsswt 2>&1 | cut -c1-300 | tee $O/nsswt.txt
PDWT_PATH=generic python tools/bench_configs.py nsswt 2>&1 | cut -c1-300 | tee -a $O/nsswt.txt
