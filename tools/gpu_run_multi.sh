#!/bin/bash
# usage: gpu_run_multi.sh N [rows]
N=$1; ROWS=${2:-100000000}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --gallery-rows $ROWS > gpurun_out/bench_n$N.log 2>&1
echo rc=$?; tail -n 3 gpurun_out/bench_n$N.log | cut -c1-2500
