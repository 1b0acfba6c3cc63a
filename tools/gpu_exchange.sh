#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/test_exchange.py > gpurun_out/exchange_n$N.log 2>&1
echo rc=$?; tail -n 5 gpurun_out/exchange_n$N.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_p2p_n$N.log 2>&1
echo rc=$?; tail -n 2 gpurun_out/bench_p2p_n$N.log | cut -c1-700
