#!/bin/bash
O=gpurun_out/exp66; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench.json; python -c "
import json; d=json.load(open('$O/bench.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value']); b=d['batched_4096x8']; print('batched', b['ms_per_step'], b['step_frac_of_peak'], json.dumps(b['level1_kernels'])); print(json.dumps(d['configs'])[:1500])"
