#!/bin/bash
O=gpurun_out/exp44; mkdir -p $O
PDWT_INV_TMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or golden or c2" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 1" "4096 4096 8" "2048 2048 64"; do
  run PDWT_INV_TMA=0
  run PDWT_INV_TMA=1
done
SHAPE="4096 4096 8"; run PDWT_INV_TMA=1 PDWT_TM=64; run PDWT_INV_TMA=1 PDWT_TM=128
