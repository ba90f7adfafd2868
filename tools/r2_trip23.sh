#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"; cut -c1-900 gpurun_out/bench_n$N.json; grep -A12 Traceback gpurun_out/bench_n$N.err | head -30
echo "=== encoder tests (split-K conv_out) on GPU 0"; timeout 900 python -m pytest tests/ -q -m gpu --timeout 600 -k "encoder or mel or smoke or tiny" 2>&1 | tail -3
