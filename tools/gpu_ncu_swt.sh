#!/bin/bash
# ncu --set full + per-instruction stall samples of the streaming SWT inverse on C3 (profiles/r02b_ncu_swt_inv_stream_summary.txt)
O=gpurun_out/ncu_swt; mkdir -p $O
PDWT_SWT_CW=256 ncu --set full --clock-control none --import-source on -k regex:"k_swt_inv" -s 4 -c 4 -o $O/ncu_swt python tools/prof_swt_once.py > $O/ncu.log 2>&1
ncu -i $O/ncu_swt.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > $O/ncu_swt_summary.txt
for k in 0 3; do
  ncu -i $O/ncu_swt.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > $O/src_$k.csv 2>/dev/null
  python tools/ncu_source_stalls.py $O/src_$k.csv 40 > $O/stalls_$k.txt; rm -f $O/src_$k.csv
done
rm -f $O/*.ncu-rep
cat $O/ncu_swt_summary.txt
