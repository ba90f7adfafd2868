"""Aggregate the `[gemm_tc time]` lines of an ASRB_GEMM_TIME=1 run (stdin): per shape count, mean us, issued-MMA TFLOP/s."""
import re, sys, collections
pat = re.compile(r"\[gemm_tc time\] M=(\d+) N=(\d+) K=(\d+) epi=(\d+) ([\d.]+) us")
acc = collections.OrderedDict()
for line in sys.stdin:
    m = pat.search(line)
    if m:
        k = tuple(int(x) for x in m.groups()[:4]); acc.setdefault(k, []).append(float(m.group(5)))
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tot = 0.0
for (M, N, K, e), v in acc.items():
    v2 = v[len(v) // 2:]                       # second half = warm pass
    us = sum(v2) / len(v2); tot += sum(v2)
    print(f"M={M:6d} N={N:5d} K={K:5d} epi={e} x{len(v2):3d}  {us:8.1f} us  {2.0 * M * N * K * planes / us * 1e-6:7.1f} TFLOP/s issued")
print(f"sum (warm half) {tot / 1000:.2f} ms")
