#!/bin/bash
mkdir -p gpurun_out
for th in 16 32 64 128; do
  echo "== PDWT_TH=$th"
  PDWT_TH=$th timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:v['avg_us'] for k,v in d['kernels'].items() if 'fwd' in k})"
done
echo "== default"; timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:v['avg_us'] for k,v in d['kernels'].items() if 'fwd' in k})"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
