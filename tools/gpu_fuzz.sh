#!/bin/bash
# randomised parity sweeps against the CPU oracle: all transform families at the default sizes, then the separable 2-D DWT
# and SWT families at sizes that span several column tiles / row chunks, then the TMA-fed inverse forced on
O=gpurun_out/fuzz; mkdir -p $O
timeout 900 python tools/fuzz_gpu.py 500 2024 > $O/fuzz_all.log 2>&1; echo "rc=$?" >> $O/fuzz_all.log; tail -2 $O/fuzz_all.log
PDWT_FUZZ_MODE=dwt2 PDWT_FUZZ_HI=1600 timeout 900 python tools/fuzz_gpu.py 120 5 > $O/fuzz_dwt2_large.log 2>&1; echo "rc=$?" >> $O/fuzz_dwt2_large.log; tail -2 $O/fuzz_dwt2_large.log
PDWT_INV_TMA=1 PDWT_FUZZ_MODE=dwt2 PDWT_FUZZ_HI=1600 timeout 900 python tools/fuzz_gpu.py 120 6 > $O/fuzz_dwt2_tma.log 2>&1; echo "rc=$?" >> $O/fuzz_dwt2_tma.log; tail -2 $O/fuzz_dwt2_tma.log
PDWT_FUZZ_MODE=swt2 PDWT_FUZZ_HI=1600 timeout 900 python tools/fuzz_gpu.py 120 7 > $O/fuzz_swt2_large.log 2>&1; echo "rc=$?" >> $O/fuzz_swt2_large.log; tail -2 $O/fuzz_swt2_large.log
PDWT_FUZZ_MODE=nsswt2 PDWT_FUZZ_HI=400 timeout 900 python tools/fuzz_gpu.py 60 8 > $O/fuzz_nsswt2.log 2>&1; echo "rc=$?" >> $O/fuzz_nsswt2.log; tail -2 $O/fuzz_nsswt2.log
PDWT_FUZZ_MODE=dwt1 timeout 900 python tools/fuzz_gpu.py 150 9 > $O/fuzz_dwt1.log 2>&1; echo "rc=$?" >> $O/fuzz_dwt1.log; tail -2 $O/fuzz_dwt1.log
