#!/bin/bash
O=gpurun_out/exp23; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
run() { timeout 120 env "$@" python tools/time_step.py $SHAPE 2>&1 | tail -1 | tee -a $O/times.txt; }
for SHAPE in "4096 4096 1" "4096 4096 8"; do
  run PDWT_MULTI=0
  run PDWT_LAG=2
done
PDWT_MULTI=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 6 -c 6 -o $O/ncu_b16 python tools/prof_batch.py > $O/ncu_b16.log 2>&1
ncu -i $O/ncu_b16.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_b16_summary.txt
for k in 0 5; do
  ncu -i $O/ncu_b16.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > $O/src_$k.csv 2>/dev/null
  python tools/ncu_source_stalls.py $O/src_$k.csv 40 > $O/stalls_$k.txt
done
rm -f $O/ncu_b16.ncu-rep
cat $O/ncu_b16_summary.txt | cut -c1-600
