#!/bin/bash
# SWT parity tests, randomised SWT cases up to 1400^2 against the oracle, per-level timing of C3, CTA-width sweep of the streaming inverse
O=gpurun_out/swt_check; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "swt or w1d2" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
PDWT_FUZZ_MODE=swt2 PDWT_FUZZ_HI=1400 timeout 900 python tools/fuzz_gpu.py 120 77 > $O/fuzz_swt2.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz_swt2.log; tail -3 $O/fuzz_swt2.log
python tools/prof_swt.py 2>&1 | tee $O/stream.txt
python tools/bench_configs.py c3 2>&1 | tee $O/c3.txt
for cw in 256 384 512; do echo "CW=$cw"; PDWT_SWT_CW=$cw python tools/prof_swt.py 2>&1 | tee $O/cw$cw.txt; done
