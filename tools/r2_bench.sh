#!/bin/bash
# both bench arms exactly as the driver launches them (default flags), with wall-clock seconds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $? wall $(( $(date +%s) - t0 )) s"; cat gpurun_out/bench_ref.json
t0=$(date +%s); echo "=== bench (default flags)"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? wall $(( $(date +%s) - t0 )) s"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
