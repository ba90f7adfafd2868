#!/bin/bash
# per-kernel durations of one transcription (ncu launch list, serialised / cold-cache: shares, not absolute times)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    v = v / 1e3 if r[ui] == "ns" else v
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{t:10.1f} us {n:5d}x avg {t/n:9.2f} us {100*t/tot:5.1f}%  {k}")
PY
