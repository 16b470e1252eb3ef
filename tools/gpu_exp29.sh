#!/bin/bash
O=gpurun_out/exp29; mkdir -p $O
python tools/timeline_multi.py 4096 8 3 2>&1 | tee $O/tl_multi_b8.txt
python tools/timeline_multi.py 4096 1 3 2>&1 | tee $O/tl_multi_b1.txt
python tools/timeline_multi.py 2048 64 3 2>&1 | tee $O/tl_multi_2048b64.txt
