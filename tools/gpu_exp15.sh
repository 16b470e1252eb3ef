#!/bin/bash
# final check of HEAD: smoke, the bench line with the concurrent-stream extra, the reference arm
O=gpurun_out/exp15; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 200 python bench.py --steps 40 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/exp15/bench_ours.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('concurrent_streams'), d.get('cpu_baseline',{}).get('value'))
PY
tail -2 $O/bench_ours.err
