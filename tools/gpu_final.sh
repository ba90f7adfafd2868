#!/bin/bash
# full validation + evidence for the shipped configuration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_all.json 2>/dev/null
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cat gpurun_out/bench_ref.json
echo "=== decode step roofline at batch 1 / 8 / 16"; timeout 600 python tools/batch_decode_roofline.py 1 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; cat gpurun_out/batch_decode.json | head -80; tail -3 gpurun_out/batch_decode.err
echo "=== encoder roofline at batch 8"; timeout 600 python tools/encoder_roofline.py 8 > gpurun_out/enc_b8.json 2> gpurun_out/enc_b8.err; echo "exit $?"; cat gpurun_out/enc_b8.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
echo "=== ncu full: decode step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 5 -c 1 -o gpurun_out/prof_mega -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
