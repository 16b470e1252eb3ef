#!/bin/bash
O=gpurun_out/exp42; mkdir -p $O
ncu --set full --clock-control none -k regex:k_rows -s 2 -c 2 -o $O/ncu_rows python tools/prof_rows.py > $O/ncu.log 2>&1
ncu -i $O/ncu_rows.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py | tee $O/summary.txt | cut -c1-700
ncu -i $O/ncu_rows.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for k in ('sm__maximum_warps_per_active_cycle_pct','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','launch__shared_mem_config_size','launch__shared_mem_per_block_dynamic','l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum'):
    if k in h: print(k, [r[h.index(k)] for r in rows[2:]])
"
rm -f $O/*.ncu-rep
