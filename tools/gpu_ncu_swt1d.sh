#!/bin/bash
# ncu --set full of the all-levels 1-D SWT kernels (profiles/r02d_ncu_swt1d_summary.txt)
O=gpurun_out/ncu_swt1d; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_rows_swt" -s 2 -c 2 -o $O/ncu python tools/prof_swt1d_once.py > $O/ncu.log 2>&1
ncu -i $O/ncu.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/summary.txt
rm -f $O/*.ncu-rep
cat $O/summary.txt
