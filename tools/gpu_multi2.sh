#!/bin/bash
# 2-GPU check after the host-publication change: NCCL sharding (norms included), bench at N=2 (c2 and c5)
mkdir -p gpurun_out/r01m
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_nccl_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu 2> gpurun_out/r01m/bench_n2.err | tee gpurun_out/r01m/bench_n2.json | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload c5 --no-cpu 2> gpurun_out/r01m/bench_c5_n2.err | tee gpurun_out/r01m/bench_c5_n2.json | cut -c1-330
tail -2 gpurun_out/r01m/bench_c5_n2.err
