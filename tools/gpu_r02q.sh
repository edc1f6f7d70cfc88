#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -s -k "cta_pair or fused_head" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02q_pytest.log
cat gpurun_out/r02q_pytest.log
for pair in 1 0; do
  CIAOSR_HEAD_PAIR=$pair timeout 300 python bench.py --steps 10 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02q_bench_pair$pair.json 2> gpurun_out/r02q_bench_pair$pair.err
  tail -2 gpurun_out/r02q_bench_pair$pair.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02q_bench_pair$pair.json'))
    print('pair=$pair', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'], d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
PY
done
