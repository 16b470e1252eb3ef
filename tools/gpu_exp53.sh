#!/bin/bash
O=gpurun_out/exp53; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "swt" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python tools/prof_swt.py 2>&1 | tee $O/stream.txt
for ch in 32 48 96; do echo "CH=$ch"; PDWT_SWT_CH=$ch python tools/prof_swt.py 2>&1 | tee $O/ch$ch.txt; done
for cw in 256 288 384 512; do echo "CW=$cw"; PDWT_SWT_CW=$cw python tools/prof_swt.py 2>&1 | tee $O/cw$cw.txt; done
echo "CW=256 CH=32"; PDWT_SWT_CW=256 PDWT_SWT_CH=32 python tools/prof_swt.py 2>&1 | tee $O/cw256ch32.txt
