#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${NGPU:-2}
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "=== bench reference arm (rank0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "exit $?"; cat gpurun_out/bench_ref_n$N.json; tail -3 gpurun_out/bench_ref_n$N.err
