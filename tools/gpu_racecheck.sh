#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels that exchange data through double-buffered shared
# rows: streaming SWT inverse, the all-levels 1-D kernels, the tiled SWT forward.  (The TMA kernels' mbarrier protocol
# is outside racecheck's model: it reports false positives there.)
O=gpurun_out/racecheck; mkdir -p $O
PDWT_INV_TMA=0 timeout 500 compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_against_oracle and stream and (w1d2 or w1d1 or w0d1)" > $O/racecheck.log 2>&1; echo "rc=$?" >> $O/racecheck.log
grep -c "Race reported\|hazard" $O/racecheck.log; grep "Race reported\|hazard" $O/racecheck.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12; tail -4 $O/racecheck.log
