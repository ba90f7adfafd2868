#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --workload b1 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"; cut -c1-1200 gpurun_out/bench_n2.json
