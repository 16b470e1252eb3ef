#!/bin/bash
# bit-exactness stress of the TMA-fed inverse level kernel against the default one, then bench.py with and without it
O=gpurun_out/tma_inv_check; mkdir -p $O
for cfg in "1 4096 94" "2 4096 46" "2 4096 64" "2 4096 128" "2 4096 94" "2 4096 32" "4 2048 94" "8 2048 94" "8 4096 0"; do set -- $cfg
  if [ "$3" = "0" ]; then B=$1 N=$2 REPS=6 python tools/tma_inv_stress.py 2>&1 | tail -1; else B=$1 N=$2 REPS=6 PDWT_TM=$3 python tools/tma_inv_stress.py 2>&1 | tail -1; fi
done | tee $O/stress.txt
B=4 N=2048 REPS=4 WNAME=sym8 LEVELS=3 python tools/tma_inv_stress.py 2>&1 | tail -1 | tee -a $O/stress.txt
B=3 N=4096 REPS=4 WNAME=db6 LEVELS=3 python tools/tma_inv_stress.py 2>&1 | tail -1 | tee -a $O/stress.txt
echo "--- bench default"; python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step']); print('batched',json.dumps(d.get('batched_4096x8'))[:600])" | tee $O/bench_default.txt
echo "--- bench PDWT_INV_TMA=1"; PDWT_INV_TMA=1 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step']); print('batched',json.dumps(d.get('batched_4096x8'))[:600])" | tee $O/bench_tma.txt
