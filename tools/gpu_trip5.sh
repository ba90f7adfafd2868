#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_pytest() {  # name, timeout, env..., -- pytest args
  name=$1; to=$2; shift 2
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== pytest $name (${envs[*]})"
  env "${envs[@]}" timeout $to python -m pytest tests/ -q -m gpu --timeout 600 "$@" > gpurun_out/pytest_$name.log 2>&1
  rc=$?; echo "pytest $name exit $rc"; tail -6 gpurun_out/pytest_$name.log
  cp gpurun_out/parity_report.json gpurun_out/parity_$name.json 2>/dev/null
  return $rc
}
run_pytest all 1200 X=1 -- ; ALL=$?
if [ $ALL -ne 0 ]; then
  run_pytest attn_simt 1200 ASRB_ATTN=simt -- ; [ $? -eq 0 ] && export ASRB_ATTN=simt
fi
grep -E "full_" gpurun_out/parity_report.json
echo "=== mega timeline"; timeout 300 python tools/mega_timeline.py > gpurun_out/mega_timeline.txt 2>&1; echo "exit $?"; tail -22 gpurun_out/mega_timeline.txt
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
