#!/bin/bash
O=gpurun_out/exp34; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err; echo "rc=$?"; tail -3 $O/bench_ours.err
python - <<'P'
import json
d=json.load(open('gpurun_out/exp34/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['step_algorithmic_gbs'])
print('batched', d['batched_4096x8']['ms_per_step'], d['batched_4096x8']['step_frac_of_peak'])
for k,v in d.get('configs',{}).items():
    print(k, json.dumps(v)[:600])
print(d.get('cpu_baseline'))
P
