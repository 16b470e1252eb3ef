#!/bin/bash
O=gpurun_out/exp61; mkdir -p $O
cp pdwt_b200/libpdwt_b200.so /tmp/lib_default.so
for d in 1 2; do cp pdwt_b200/_alt/lib_d$d.so pdwt_b200/libpdwt_b200.so; echo "DIAG=$d"; python tools/prof_swt.py 2>&1 | tee $O/d$d.txt; done
cp /tmp/lib_default.so pdwt_b200/libpdwt_b200.so
