#!/bin/bash
O=gpurun_out/exp43; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or golden or c2" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 1" "4096 4096 8" "2048 2048 64"; do
  run PDWT_X=0
done
