#!/bin/bash
O=gpurun_out/exp31; mkdir -p $O
run() { timeout 120 env "$@" 2>&1 | tail -2 | tee -a $O/times.txt; }
run PDWT_MULTI=0 python tools/time_streams.py 4096 4096 8 8
run PDWT_MULTI=0 python tools/time_streams.py 4096 4096 8 4
run PDWT_MULTI=0 python tools/time_streams.py 4096 4096 8 2
run PDWT_MULTI=0 PDWT_LOWOCC=1 python tools/time_streams.py 4096 4096 8 4
run PDWT_MULTI=0 PDWT_LOWOCC=0 python tools/time_streams.py 4096 4096 8 8
run PDWT_MULTI=0 PDWT_PDL=0 python tools/time_streams.py 4096 4096 8 8
run PDWT_MULTI=0 python tools/time_streams.py 2048 2048 64 8
run PDWT_MULTI=0 python tools/time_streams.py 2048 2048 64 16
