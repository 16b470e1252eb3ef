#!/bin/bash
# run on an N-GPU box: NCCL sharding check, then bench at N (c2 per GPU and the batched c5 workload)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_nccl_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload c5 2> gpurun_out/bench_c5_n$N.err | tee gpurun_out/bench_c5_n$N.json | cut -c1-330
python bench.py --steps 10 --warmup 3 --workload c5 --no-cpu 2> gpurun_out/bench_c5_n1.err | tee gpurun_out/bench_c5_n1.json | cut -c1-330
tail -2 gpurun_out/bench_c5_n$N.err
