#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu quick"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 -k "not full_size" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_gpu_quick.log
for shape in 3120,2688,896 3120,896,3584 49920,480,4608; do
  echo "=== shape $shape"
  ASRB_GEMM_DEBUG=$shape PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2>&1 | grep -v "^ *\"\(batch\|mma\)" | tail -14
done
echo "=== encoder roofline B=8 all planes"; timeout 600 python tools/encoder_roofline.py 8 2>&1 | grep "encoder_ms\|prefill_ms"
echo "=== bench b1"; timeout 900 python bench.py --workload b1 --steps 3 --warmup 3 2>&1 | tail -1
