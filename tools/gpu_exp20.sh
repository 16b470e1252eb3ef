#!/bin/bash
# r02 exp20: compact forward loop (shifting accumulator file) + producer poll interval
O=gpurun_out/exp20; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
for ns in 40 200 1000; do
  PDWT_POLL_NS=$ns python tools/time_step.py 4096 4096 1
  PDWT_POLL_NS=$ns python tools/time_step.py 4096 4096 8
done 2>&1 | tee $O/times.txt
python tools/time_step.py 2048 2048 64 2>&1 | tee -a $O/times.txt
PDWT_LOWOCC=0 python tools/time_step.py 4096 4096 1 2>&1 | tee -a $O/times.txt
PDWT_LOWOCC=1 python tools/time_step.py 4096 4096 8 2>&1 | tee -a $O/times.txt
PDWT_TH=128 python tools/time_step.py 4096 4096 8 2>&1 | tee -a $O/times.txt
PDWT_TH=64 python tools/time_step.py 4096 4096 8 2>&1 | tee -a $O/times.txt
