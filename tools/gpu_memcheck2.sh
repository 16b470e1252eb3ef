#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 2: streaming SWT inverse, TMA-fed inverse (forced), SWT forward
O=gpurun_out/memcheck2; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w1d2 or tma or swt" > $O/memcheck_swt_tma.log 2>&1; echo "rc=$?" >> $O/memcheck_swt_tma.log; tail -4 $O/memcheck_swt_tma.log
PDWT_INV_TMA=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_against_oracle and stream and (db7 or sym8 or db6 or db9)" > $O/memcheck_tma_forced.log 2>&1; echo "rc=$?" >> $O/memcheck_tma_forced.log; tail -4 $O/memcheck_tma_forced.log
PDWT_FUZZ_MODE=swt2 PDWT_FUZZ_HI=500 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/fuzz_gpu.py 40 99 > $O/memcheck_fuzz_swt.log 2>&1; echo "rc=$?" >> $O/memcheck_fuzz_swt.log; tail -3 $O/memcheck_fuzz_swt.log
