#!/bin/bash
O=gpurun_out/exp60; mkdir -p $O
cp pdwt_b200/libpdwt_b200.so /tmp/lib_default.so
echo "default (FMUL2 + FFMA2)"; python tools/prof_swt.py 2>&1 | tee $O/mul.txt
cp pdwt_b200/_alt/lib_fma.so pdwt_b200/libpdwt_b200.so; echo "FFMA2 + FFMA2"; python tools/prof_swt.py 2>&1 | tee $O/fma.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "swt" > $O/pytest_fma.log 2>&1; tail -2 $O/pytest_fma.log
cp /tmp/lib_default.so pdwt_b200/libpdwt_b200.so
