#!/bin/bash
# N-split issue schedule of the CTA-pair kernel: parity, bench (2x2: schedule x row threads per row), timelines
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "cta_pair or fused_head" 2>&1 | grep -v "^$" | tail -14 > gpurun_out/r02y_pair_test.log
cat gpurun_out/r02y_pair_test.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -6 > gpurun_out/r02y_pytest.log
cat gpurun_out/r02y_pytest.log
for ns in 1 0; do for parts in 4 2; do
  CIAOSR_HEAD_NSPLIT=$ns CIAOSR_HEAD_ROWPARTS=$parts timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02y_bench_ns${ns}_parts$parts.json 2> gpurun_out/r02y_bench_ns${ns}_parts$parts.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02y_bench_ns${ns}_parts$parts.json'))
    print('nsplit=$ns parts=$parts', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done; done
for parts in 2 4; do
  CIAOSR_HEAD_ROWPARTS=$parts CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_trace.so timeout 300 python tools/trace_pair.py > gpurun_out/r02y_trace_parts$parts.txt 2> gpurun_out/r02y_trace_parts$parts.err
done
