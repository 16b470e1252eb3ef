#!/bin/bash
# 2-GPU (or N-GPU: NG=4 bash tools/gpu_multi2.sh) check: Layer C sharding over NCCL, bench at N with the sharded C5 block
NG=${NG:-2}; O=gpurun_out/multi$NG; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/sharded_nccl_check.py > $O/sharded_nccl_check.txt 2>&1; grep -v "^W\|^\*\*\*" $O/sharded_nccl_check.txt | tail -3
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu > $O/bench_n$NG.json 2> $O/bench_n$NG.err; echo "bench rc=$?"
grep -c "NCCL INFO" $O/bench_n$NG.json $O/bench_n$NG.err; grep -m3 "NVLS\|via P2P\|Channel 00" $O/bench_n$NG.err $O/bench_n$NG.json | cut -c1-200
grep '^{' $O/bench_n$NG.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print(json.dumps(d.get('c5_sharded'),indent=1))
"
tail -3 $O/bench_n$NG.err | cut -c1-300
