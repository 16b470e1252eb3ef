#!/bin/bash
O=gpurun_out/exp65; mkdir -p $O
python tools/time_inv_variants.py 2>&1 | tee $O/variants.txt
