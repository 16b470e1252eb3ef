#!/bin/bash
O=gpurun_out/exp35; mkdir -p $O
python tests/golden/make_golden_custom2d.py 2>&1 | tail -5
cp gpurun_out/golden/ns*2_custom_*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_reference_fullsize.py -m gpu -x -q -k "custom" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -15 $O/pytest.log
for v in 0 1; do PDWT_NS_FWD_CONST=$v python tools/bench_configs.py c4 c4db2 2>&1 | cut -c1-260; done | tee $O/ns_const.txt
