#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02k_pytest.log
cat gpurun_out/r02k_pytest.log
timeout 300 python tools/time_trunk_ops.py > gpurun_out/r02k_trunk_ops.json 2>&1; cat gpurun_out/r02k_trunk_ops.json | tr -d '\n ' ; echo
timeout 300 python tools/time_csattn.py > gpurun_out/r02k_time_csattn.jsonl 2>&1; head -3 gpurun_out/r02k_time_csattn.jsonl
for fused in 1 0; do
  CIAOSR_HEAD_FUSED=$fused timeout 600 python bench.py --steps 10 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02k_bench_fused$fused.json 2> gpurun_out/r02k_bench_fused$fused.err
  tail -2 gpurun_out/r02k_bench_fused$fused.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02k_bench_fused$fused.json'))
    print('fused=$fused', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'], d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
PY
done
