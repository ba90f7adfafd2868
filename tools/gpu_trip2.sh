#!/bin/bash
# GPU visit 2: validate the fused decode step and the tcgen05 GEMM separately (each under its own timeout,
# a deadlock in one must not cost the others), then bench the best working configuration.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_pytest() {  # name, timeout, env..., -- pytest args
  name=$1; to=$2; shift 2
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== pytest $name (${envs[*]})"
  env "${envs[@]}" timeout $to python -m pytest tests/ -q -m gpu --timeout 600 "$@" > gpurun_out/pytest_$name.log 2>&1
  rc=$?; echo "pytest $name exit $rc"; tail -15 gpurun_out/pytest_$name.log
  cp gpurun_out/parity_report.json gpurun_out/parity_$name.json 2>/dev/null
  return $rc
}
run_pytest base 900 ASRB_GEMM=simt ASRB_DECODE=phases -- -k "not mega and not full_size"; BASE=$?
run_pytest mega 420 ASRB_GEMM=simt ASRB_DECODE=mega -- -k "ids or eos"; MEGA=$?
run_pytest tc 420 ASRB_GEMM=tc ASRB_DECODE=phases -- -k "encoder or prefill or (ids and phases)"; TC=$?
ARGS=""
[ $MEGA -ne 0 ] && ARGS="$ARGS --decode phases"
[ $TC -ne 0 ] && ARGS="$ARGS --gemm simt"
echo "=== full-size test with: $ARGS"
D=mega; [ $MEGA -ne 0 ] && D=phases; GM=tc; [ $TC -ne 0 ] && GM=simt
run_pytest full 600 ASRB_GEMM=$GM ASRB_DECODE=$D -- -k "full_size"
echo "=== bench $ARGS"; timeout 900 python bench.py --steps 3 --warmup 3 $ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline $ARGS > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
