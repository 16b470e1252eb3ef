#!/bin/bash
O=gpurun_out/exp26; mkdir -p $O
PDWT_MARGIN=60 python tools/timeline_multi.py 4096 8 3 2>&1 | tee $O/tl_multi_b8.txt
PDWT_MARGIN=60 python tools/timeline_multi.py 4096 1 3 2>&1 | tee $O/tl_multi_b1.txt
PDWT_MARGIN=60 python tools/timeline_multi.py 2048 16 3 2>&1 | tee $O/tl_multi_2048b16.txt
