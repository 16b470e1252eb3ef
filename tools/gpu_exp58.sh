#!/bin/bash
O=gpurun_out/exp58; mkdir -p $O
cp pdwt_b200/libpdwt_b200.so /tmp/lib_default.so
echo "NL=2 (default)"; python tools/prof_swt.py 2>&1 | tee $O/nl2.txt
for nl in 1 4; do cp pdwt_b200/_alt/lib_nl$nl.so pdwt_b200/libpdwt_b200.so; echo "NL=$nl"; python tools/prof_swt.py 2>&1 | tee $O/nl$nl.txt; done
cp /tmp/lib_default.so pdwt_b200/libpdwt_b200.so
