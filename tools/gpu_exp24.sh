#!/bin/bash
O=gpurun_out/exp24; mkdir -p $O
python tools/timeline_batch.py 4096 8 2>&1 | tee $O/tl_b8.txt
PDWT_TH=128 python tools/timeline_batch.py 4096 8 2>&1 | tee $O/tl_b8_th128.txt
PDWT_LOWOCC=1 python tools/timeline_batch.py 4096 8 2>&1 | tee $O/tl_b8_low.txt
