#!/bin/bash
# 16 row warps (four row threads per row) in the CTA-pair kernel: parity + bench against the two-thread form
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "cta_pair" 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02v_pair_test.log
cat gpurun_out/r02v_pair_test.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r02v_pytest.log
cat gpurun_out/r02v_pytest.log
for parts in 4 2; do
  CIAOSR_HEAD_ROWPARTS=$parts timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02v_bench_parts$parts.json 2> gpurun_out/r02v_bench_parts$parts.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02v_bench_parts$parts.json'))
    print('parts=$parts', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done
