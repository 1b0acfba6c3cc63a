#!/bin/bash
# usage: gpu_multi.sh N : exchange correctness test + bench lines (1M / 10M / 100M rows) at N GPUs, each with the
# in-run parity check; JSON lines land in gpurun_out/r02_bench_{rows}_n{N}.json
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/test_exchange.py > gpurun_out/r02_exchange_n$N.log 2>&1
echo "exchange rc=$?"; tail -n 3 gpurun_out/r02_exchange_n$N.log | cut -c1-300
[ $N = 2 ] && { timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2; }
p=29520
for rows in 1000000 10000000 100000000; do
  p=$((p+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --steps 5 --warmup 3 --gallery-rows $rows > gpurun_out/r02_bench_${rows}_n$N.json 2> gpurun_out/r02_bench_${rows}_n$N.err
  echo "rows=$rows rc=$?"; tail -n 1 gpurun_out/r02_bench_${rows}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
pc=d['parity_check']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','status_ok')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clk', d['clocks']['sm_mhz'], 'pc', pc['ok'], pc['missed_rows'], pc['max_rank_gap'], pc['ranks'])"
done
