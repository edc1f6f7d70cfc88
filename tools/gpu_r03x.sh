#!/bin/bash
# synccheck of the head kernels on small cases + full GPU tests
set -u
mkdir -p gpurun_out
for k in fused_head cta_pair; do
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$k" > gpurun_out/r03x_synccheck_$k.log 2>&1
echo "== synccheck $k"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r03x_synccheck_$k.log | head -4
grep -A6 "Barrier error" gpurun_out/r03x_synccheck_$k.log | grep "Device Frame" | sort | uniq -c | head -5 | cut -c1-200
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -2
