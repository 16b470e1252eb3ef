#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/prof_seq.py 2>&1 | tee gpurun_out/seq.txt
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['sync']['value'])
print(d['roofline']); print(d['level1_back_to_back']); print(d['kernels']); print(d.get('cpu_baseline'))
P
tail -3 gpurun_out/bench_ours.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload c2b8 --no-cpu > gpurun_out/bench_c2b8.json 2> gpurun_out/bench_c2b8.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_c2b8.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['step_algorithmic_gbs'], d['roofline'], d['level1_back_to_back'])
P
