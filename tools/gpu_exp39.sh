#!/bin/bash
O=gpurun_out/exp39; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
SHAPE="4095 4097 1"; run PDWT_X=0
SHAPE="4096 4100 1"; run PDWT_X=0
SHAPE="4095 4097 8"; run PDWT_X=0
