#!/bin/bash
O=gpurun_out/exp40; mkdir -p $O
for r in 4 6 8; do
  PDWT_BENCH_ROTATE=$r timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rotate', $r, 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'sync', d['e2e']['sync']['value'])"
done | tee $O/e2e.txt
