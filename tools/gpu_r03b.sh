#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair or fused_head or full" 2>&1 | grep -v "^$" | tail -4
for ns in 0 1 0; do
  CIAOSR_HEAD_NSPLIT=$ns CIAOSR_HEAD_ROWPARTS=2 timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/${TAG}_bench_ns${ns}.json 2> gpurun_out/${TAG}_bench_ns${ns}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_ns${ns}.json'))
    print('nsplit=$ns', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done
