#!/bin/bash
O=gpurun_out/exp51; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "swt" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
python tools/prof_swt.py 2>&1 | tee $O/stream.txt
PDWT_SWT_INV_STREAM=0 python tools/prof_swt.py 2>&1 | tee $O/tiled.txt
for ch in 32 128; do echo "CH=$ch"; PDWT_SWT_CH=$ch python tools/prof_swt.py 2>&1 | tee $O/ch$ch.txt; done
for cw in 288 416 640; do echo "CW=$cw"; PDWT_SWT_CW=$cw python tools/prof_swt.py 2>&1 | tee $O/cw$cw.txt; done
