#!/bin/bash
O=gpurun_out/exp55; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "swt" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python tools/prof_swt.py 2>&1 | tee $O/stream.txt
for pf in 0 4 16; do echo "PF=$pf"; PDWT_SWT_PF=$pf python tools/prof_swt.py 2>&1 | tee $O/pf$pf.txt; done
