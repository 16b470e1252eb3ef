#!/bin/bash
O=gpurun_out/exp41; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_prox_ops.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -6 $O/pytest.log
python tools/bench_configs.py dwt1d 2>&1 | cut -c1-300 | tee $O/dwt1d.txt
