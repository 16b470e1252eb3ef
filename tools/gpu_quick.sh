#!/bin/bash
# one short call: the GPU test suite, the default bench line (value, e2e, batched block) and a randomised 2-D parity sweep
O=gpurun_out/quick; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python bench.py --steps 40 --warmup 5 2>/dev/null | tail -1 > $O/bench.json; python -c "
import json; d=json.load(open('$O/bench.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value']); b=d['batched_4096x8']; print('batched', b['ms_per_step'], b['step_frac_of_peak'], json.dumps(b['level1_kernels']))"
PDWT_FUZZ_MODE=dwt2 PDWT_FUZZ_HI=1200 timeout 600 python tools/fuzz_gpu.py 150 31 2>&1 | tail -1
